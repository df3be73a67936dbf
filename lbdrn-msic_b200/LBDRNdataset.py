"""Per-pixel dataset of LBDRN -- drop-in for the reference's LBDRNdataset.py (same public names:
`merge_tiles`, `split_image`, `write_tiff_with_gdal`, `process`, `LBDRNDataset`).

What changed: `LBDRNDataset` no longer materialises the (H*W, dim_in) float32 feature matrix on the host
(reference LBDRNdataset.py:104-143: 400 B/pixel, the encoder's dominant host cost).  The scene is split into integer
MSB/LSB planes on the GPU (`lbdrn_fused.DeviceScene`) and the fused kernels sample neighbourhoods from those planes.
`process()` and `__getitem__` remain as cold, host-side compatibility paths producing the reference's values.
Raster I/O still goes through GDAL, unchanged (the package is provided by the environment).
"""
import os

import numpy as np
import torch
from osgeo import gdal
from torch.utils.data import Dataset

import constants

gdal.UseExceptions()


# ---- host raster helpers (reference LBDRNdataset.py:12-89; plain GDAL calls) --------------------------------------
def write_tiff_with_gdal(output_path, array):
    """Write a CHW numpy array as a GeoTIFF, one band per plane."""
    kinds = {np.uint8: gdal.GDT_Byte, np.uint16: gdal.GDT_UInt16, np.float32: gdal.GDT_Float32,
             np.float64: gdal.GDT_Float64}
    if array.dtype.type not in kinds:
        raise ValueError("Unsupported data type in this function")
    bands, rows, cols = array.shape
    ds = gdal.GetDriverByName('GTiff').Create(output_path, cols, rows, bands, kinds[array.dtype.type])
    for b in range(bands):
        ds.GetRasterBand(b + 1).WriteArray(array[b])
    ds.FlushCache()
    ds = None


def _tile_windows(width, height, split_ratio):
    """(i, j, x, y, w, h) of the sr x sr tiling; the last row/column absorbs the remainder."""
    tw, th = width // split_ratio, height // split_ratio
    for i in range(split_ratio):
        for j in range(split_ratio):
            x, y = j * tw, i * th
            yield (i, j, x, y, width - x if j + 1 == split_ratio else tw, height - y if i + 1 == split_ratio else th)


def split_image(input_file, output_dir, split_ratio):
    src = gdal.Open(input_file)
    if src is None:
        raise ValueError(f"Failed to open {input_file}.")
    for i, j, x, y, w, h in _tile_windows(src.RasterXSize, src.RasterYSize, split_ratio):
        gdal.Translate(os.path.join(output_dir, f"tile_{i}_{j}.tif"), src, srcWin=[x, y, w, h])
        print(f"Tile {i}_{j} created, shape: {w}x{h}")
    src = None


def merge_tiles(input_dir, output_file, split_ratio, width, height):
    first = gdal.Open(os.path.join(input_dir, "tile_0_0_recon.tif"))
    out = gdal.GetDriverByName('GTiff').Create(output_file, width, height, first.RasterCount,
                                               first.GetRasterBand(1).DataType)
    if out is None:
        raise ValueError(f"Failed to create {output_file}.")
    for i, j, x, y, w, h in _tile_windows(width, height, split_ratio):
        path = os.path.join(input_dir, f"tile_{i}_{j}_recon.tif")
        tile = gdal.Open(path)
        if tile is None:
            raise ValueError(f"Failed to open {path}.")
        out.WriteArray(tile.ReadAsArray(), x, y)
        print(f"Tile {i}_{j} merged, shape: {w}x{h}")
    out.FlushCache()
    out = None


# ---- cold host path: explicit features / labels ---------------------------------------------------------------------
def _split_host(img, K):
    img = img.reshape((-1,) + img.shape[-2:])
    msb = img >> K
    lsb = (img - (msb << K)).astype(np.float32) / (2 ** K - 1)
    return (msb.astype(np.uint16) if msb.max() > 255 else msb.astype(np.uint8)), lsb


def host_features(msb, D):
    """(H*W, dim_in) float32 feature matrix with the reference's column layout, built plane by plane from shifted
    views of the reflect-padded, max-normalised MSB image (cold path; the GPU path never builds this)."""
    import lbdrn_fused
    fl = lbdrn_fused.Flags.from_constants()
    C, H, W = msb.shape
    n = 2 * D + 1
    nco = fl.num_coords()
    out = np.zeros((H, W, fl.dim_in(C, D)), dtype=np.float32)
    if fl.use_coordinates:
        tab = lbdrn_fused.coord_table(H, W, fl)
        out[:, :, :fl.tabw] = tab[:H][:, None, :]
        out[:, :, fl.tabw:nco] = tab[H:][None, :, :]
    if fl.use_colors:
        s = np.pad(msb.astype(np.float32) / msb.max(), ((0, 0), (D, D), (D, D)), mode='reflect')
        for c in range(C):
            ctr = s[c, D:D + H, D:D + W] if (fl.relative and D > 0) else 0
            for dy in range(n):
                for dx in range(n):
                    out[:, :, nco + (c * n + dy) * n + dx] = s[c, dy:dy + H, dx:dx + W] - ctr
    return out.reshape(H * W, -1)


def process(path, K, D, output_path):
    """Reference-compatible: read the raster, write the MSB base layer to `output_path`, return the explicit
    (features, labels) matrices (reference LBDRNdataset.py:92-133)."""
    msb, lsb = _split_host(gdal.Open(path).ReadAsArray(), K)
    write_tiff_with_gdal(output_path, msb)
    return host_features(msb, D), np.ascontiguousarray(lsb.transpose(1, 2, 0).reshape(-1, lsb.shape[0]))


# Scenes already uploaded by the batch scheduler (lbdrn_sched.py): absolute path -> CHW uint16 CUDA tensor.  A K sweep over
# one scene then reads and uploads the raster once; the MSB/LSB split for each K runs on the device.
PRELOADED = {}


def preload(path, device=None):
    import lbdrn_fused
    img = gdal.Open(path).ReadAsArray()
    img = img.reshape((-1,) + img.shape[-2:])
    t = torch.from_numpy(np.ascontiguousarray(img.astype(np.uint16, copy=False))).to(lbdrn_fused._device(device))
    PRELOADED[os.path.abspath(path)] = t
    return t


class LBDRNDataset(Dataset):
    """Scene resident on the GPU as MSB/LSB planes; same attributes as the reference's dataset
    (`n_pixels, n_feature, channels, n_subpixels`) and the same side effect of writing `<out>/<name>_base.tif`."""

    def __init__(self, args):
        import lbdrn_fused
        name = os.path.splitext(os.path.basename(args.path))[0]
        self.K, self.D = args.K, args.D
        self.flags = lbdrn_fused.Flags.from_constants()
        img = PRELOADED.get(os.path.abspath(args.path))
        if img is None:
            img = gdal.Open(args.path).ReadAsArray()
        self.scene = lbdrn_fused.DeviceScene.from_image(img, args.K)
        self._msb_host = self.scene.msb.cpu().numpy()
        write_tiff_with_gdal(f'{args.output_dir}/{name}_base.tif', self._msb_host)
        self.channels = self.scene.C
        self.n_pixels = self.scene.H * self.scene.W
        self.n_feature = self.flags.dim_in(self.scene.C, args.D)
        self.n_subpixels = self.n_pixels * self.channels
        self._cold = None

    def __len__(self):
        return self.n_pixels

    def __getitem__(self, idx):
        """Cold path (explicit feature row + label of one pixel); builds the host matrices on first use."""
        if self._cold is None:
            lsb = self.scene.lsb.cpu().numpy().astype(np.float32) / (2 ** self.K - 1)
            self._cold = (torch.from_numpy(host_features(self._msb_host, self.D)),
                          torch.from_numpy(np.ascontiguousarray(lsb.transpose(1, 2, 0).reshape(self.n_pixels, -1))))
        return self._cold[0][idx], self._cold[1][idx]
