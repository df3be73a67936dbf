"""TEST INFRASTRUCTURE ONLY -- stand-in for the `osgeo` (GDAL) package, which is not installed in this image.

The reference (and our drop-in CLI) does raster I/O through `osgeo.gdal` (reference LBDRNdataset.py:4,
encode.py:12, decode.py:13).  This shim implements just the calls those files make, on top of a trivial
"npy inside a .tif-named file" container, so that the reference can be executed verbatim as the oracle and
so that our CLI can be exercised end to end where GDAL is absent.  It is never imported by product code
paths: product code does `from osgeo import gdal` and gets the real GDAL when it is installed.
"""
from . import gdal  # noqa: F401
