"""Stand-in so the reference imports in this container (test infrastructure only)."""
from functools import cached_property  # noqa: F401
