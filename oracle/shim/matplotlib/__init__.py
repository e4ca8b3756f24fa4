"""Stand-in for matplotlib (test infrastructure only)."""
