"""Stand-in for the reference's `copo` package name (see copo_b200/compat/__init__.py)."""
