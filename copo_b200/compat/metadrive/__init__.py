"""Stand-in for the `metadrive` package name (see copo_b200/compat/__init__.py): the multi-agent driving environments
are this repo's batched CUDA simulator behind MetaDrive's dict API, not MetaDrive."""
__version__ = "0.2.5+copo_b200"
