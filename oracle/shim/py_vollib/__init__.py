"""Stand-in for py_vollib: only constants, exceptions and the modules the reference patches."""
