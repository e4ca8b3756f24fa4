"""Stand-in for `quimb` (test infrastructure only)."""
