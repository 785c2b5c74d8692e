"""Mirror of the reference's ``EMIT_data`` call surface for the GLT orthorectification hot path.

Unlike the reference's ``EMIT_data/__init__.py:1-11`` nothing heavy (earthaccess, xarray, GDAL)
is imported at import time; only the gather-related names of ``emit_proj.py`` and
``emit_tools.py`` are provided.
"""
from .emit_proj import NO_DATA_VALUE, glt_ortho, ortho_planes  # noqa: F401
from .emit_tools import apply_glt  # noqa: F401
from .nc_export import convert_emit_nc_to_envi, get_attr, nc_to_envi, open_any_nc, run_cmd  # noqa: F401,E402
