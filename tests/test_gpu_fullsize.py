"""GPU, BASELINE.json's full sizes: parity through size-independent properties + sampled oracle windows.

The oracle needs ~400 B/pixel, so whole-scene comparison is impossible at 8192^2; instead
  * random windows of the full-size output are compared with the oracle run on (window + margin) crops in which one
    corner pixel is set to the scene's global maximum (the normaliser is a whole-image property; the planted pixel is
    farther than D from the compared interior, so it cannot influence it);
  * stripe-sharded and host-streamed decodes must be bit-identical to the resident whole-image decode;
  * encoder-side reconstruction and decoder output are bit-identical (same kernel, same stream of weights)."""
import ctypes

import numpy as np
import pytest
import torch

from conftest import load_case, split_stream
import fpzip  # shim
import lbdrn_cabi as cabi
import lbdrn_fused as F
import lbdrn_oracle as O

pytestmark = pytest.mark.gpu


def _params(name="k5d2_train"):
    _, _, blob, _ = load_case(name)
    _, tiles = split_stream(blob)
    return np.asarray(fpzip.decompress(tiles[0][0])[0][0][0], dtype=np.float32)


def _window_check(msb_dev, out_dev, params, K, D, dim_in, bc, C, nl, n_windows=6, win=64, margin=8, seed=0):
    rng = np.random.default_rng(seed)
    Cc, H, W = msb_dev.shape
    gmax = int(msb_dev.to(torch.int32).max()) if msb_dev.dtype != torch.uint16 else int(msb_dev.view(torch.int16).to(torch.int32).max())
    p = O.unflatten_params(params, dim_in, bc, C, nl)
    bad = tot = 0
    for _ in range(n_windows):
        y0 = int(rng.integers(margin, H - win - margin))
        x0 = int(rng.integers(margin, W - win - margin))
        crop = msb_dev[:, y0 - margin:y0 + win + margin, x0 - margin:x0 + win + margin]
        crop = (crop.view(torch.int16) if crop.dtype == torch.uint16 else crop).cpu().numpy().copy()
        if crop.dtype == np.int16:
            crop = crop.view(np.uint16)
        crop[0, 0, 0] = gmax                                   # plant the global normaliser (outside the interior)
        ref = O.decode_image(crop, p, K, D)[:, margin:-margin, margin:-margin]
        got = out_dev[:, y0:y0 + win, x0:x0 + win]
        got = got.view(torch.int16).cpu().numpy().view(np.uint16)
        diff = np.abs(got.astype(np.int64) - ref.astype(np.int64))
        assert diff.max() <= 1
        bad += int((diff != 0).sum())
        tot += diff.size
    assert bad <= max(1, int(1e-4 * tot)), (bad, tot)


def test_config2_8192_decode_windows_stripes_and_stream():
    from synth_scene import make_scene_torch
    K, D = 5, 2
    img = make_scene_torch(4, 8192, 8192, 12, seed=5, device="cuda")
    scene = F.DeviceScene.from_image(img, K)
    del img
    params = _params()
    pd = torch.from_numpy(params).cuda()
    whole = F.decode_image(scene.msb, pd, K, D, 64, 2, flags=F.Flags(), return_tensor=True, base_max=scene.msb_max)
    _window_check(scene.msb, whole, params, K, D, 100, 64, 4, 2)
    # row-stripe sharding (two uneven stripes with halos) is bit-identical
    lib = cabi.load()
    out = torch.empty_like(whole)
    for r0, r1 in ((0, 3001), (3001, 8192)):
        d = cabi.make_desc(4, 8192, 8192, K, D, 64, 2, F.Flags().bits(), scene.msb_max, False, row0=r0, row1=r1)
        cabi.check(lib.lbdrn_decode(ctypes.byref(d), cabi.ptr(scene.msb), cabi.ptr(pd), None, cabi.ptr(out), cabi.stream_ptr()))
    assert torch.equal(out.view(torch.int16), whole.view(torch.int16))
    # host -> host streamed decode is bit-identical
    host = scene.msb.cpu().pin_memory()
    streamed = F.decode_image_streamed(host, params, K, D, 64, 2, flags=F.Flags())
    assert torch.equal(streamed.view(torch.int16), whole.cpu().view(torch.int16))
    # the precise fp32 path agrees with the tensor path within the bar on a 1024-row band
    d = cabi.make_desc(4, 8192, 8192, K, D, 64, 2, F.Flags().bits(), scene.msb_max, False, row0=4096, row1=5120,
                       path=cabi.PATH_PRECISE)
    cabi.check(lib.lbdrn_decode(ctypes.byref(d), cabi.ptr(scene.msb), cabi.ptr(pd), None, cabi.ptr(out), cabi.stream_ptr()))
    a = out[:, 4096:5120].view(torch.int16).to(torch.int32)
    b = whole[:, 4096:5120].view(torch.int16).to(torch.int32)
    diff = (a - b).abs()
    assert int(diff.max()) <= 1 and float((diff != 0).float().mean()) <= 1e-4


def test_config5_shape_8band_16bit_stripe_decode():
    """GF-6-shaped: 8 bands, 16-bit, K=5 (MSB up to 2047 -> uint16 planes), one 2048-row stripe of a 16384-wide scene,
    decoded as two stripes with halos; windows against the oracle."""
    from synth_scene import make_scene_torch
    K, D = 5, 2
    img = make_scene_torch(8, 2048, 16384, 16, seed=9, device="cuda")
    scene = F.DeviceScene.from_image(img, K)
    del img
    assert scene.msb.dtype == torch.uint16
    torch.manual_seed(1)
    from LBDRNmodel import LBDRNModel
    params = O.fpzip_value_map(LBDRNModel(200, 64, 8, 2).flat_params().numpy(), 16)
    pd = torch.from_numpy(params).cuda()
    whole = F.decode_image(scene.msb, pd, K, D, 64, 2, flags=F.Flags(), return_tensor=True, base_max=scene.msb_max)
    _window_check(scene.msb, whole, params, K, D, 200, 64, 8, 2, n_windows=4, win=48)
    lib = cabi.load()
    out = torch.empty_like(whole)
    for r0, r1 in ((0, 1024), (1024, 2048)):
        d = cabi.make_desc(8, 2048, 16384, K, D, 64, 2, F.Flags().bits(), scene.msb_max, True, row0=r0, row1=r1)
        cabi.check(lib.lbdrn_decode(ctypes.byref(d), cabi.ptr(scene.msb), cabi.ptr(pd), None, cabi.ptr(out), cabi.stream_ptr()))
    assert torch.equal(out.view(torch.int16), whole.view(torch.int16))


def test_config3_d3_bc256_fullsize():
    """BASELINE config 3 (D=3, bc=256) at 8192^2 on the wide tcgen05 kernel: oracle windows, stripe decode bit-identical,
    and the fp32 kernel within the bar on a band."""
    from synth_scene import make_scene_torch
    K, D = 5, 3
    img = make_scene_torch(4, 8192, 8192, 12, seed=11, device="cuda")
    scene = F.DeviceScene.from_image(img, K)
    del img
    params = _params("d3_bc256")
    pd = torch.from_numpy(params).cuda()
    lib = cabi.load()
    d = cabi.make_desc(4, 8192, 8192, K, D, 256, 2, F.Flags().bits(), scene.msb_max, False)
    assert lib.lbdrn_has_tensor_path(ctypes.byref(d)) == 1
    whole = F.decode_image(scene.msb, pd, K, D, 256, 2, flags=F.Flags(), return_tensor=True, base_max=scene.msb_max)
    _window_check(scene.msb, whole, params, K, D, 196, 256, 4, 2, n_windows=4, win=48)
    out = torch.empty_like(whole)
    for r0, r1 in ((0, 5003), (5003, 8192)):
        d = cabi.make_desc(4, 8192, 8192, K, D, 256, 2, F.Flags().bits(), scene.msb_max, False, row0=r0, row1=r1)
        cabi.check(lib.lbdrn_decode(ctypes.byref(d), cabi.ptr(scene.msb), cabi.ptr(pd), None, cabi.ptr(out), cabi.stream_ptr()))
    assert torch.equal(out.view(torch.int16), whole.view(torch.int16))
    d = cabi.make_desc(4, 8192, 8192, K, D, 256, 2, F.Flags().bits(), scene.msb_max, False, row0=2048, row1=2304,
                       path=cabi.PATH_PRECISE)
    cabi.check(lib.lbdrn_decode(ctypes.byref(d), cabi.ptr(scene.msb), cabi.ptr(pd), None, cabi.ptr(out), cabi.stream_ptr()))
    diff = (out[:, 2048:2304].view(torch.int16).to(torch.int32) - whole[:, 2048:2304].view(torch.int16).to(torch.int32)).abs()
    assert int(diff.max()) <= 1 and float((diff != 0).float().mean()) <= 1e-4


@pytest.mark.parametrize("case", ["coords_pe", "coords_pe_col"])
def test_config4_coordinates_fullsize(case):
    """BASELINE config 4 (USE_COORDINATES + EMBEDDING, without / with colours) at 8192^2 on the tensor path.  Coordinates
    are absolute, so the oracle is evaluated on full-width row bands of the real scene (not on crops)."""
    from synth_scene import make_scene_torch
    from conftest import case_flags
    meta, _, _, _ = load_case(case)
    fl, ofl = case_flags(meta, F.Flags), case_flags(meta, O.Flags)
    K, D, H, W = 5, 2, 8192, 8192
    img = make_scene_torch(4, H, W, 12, seed=13, device="cuda")
    scene = F.DeviceScene.from_image(img, K)
    del img
    params = _params(case)
    whole = F.decode_image(scene.msb, torch.from_numpy(params).cuda(), K, D, 64, 2, flags=fl, return_tensor=True,
                           base_max=scene.msb_max, path="tensor")
    msb = scene.msb.cpu().numpy()
    p = O.unflatten_params(params, ofl.dim_in(4, D), 64, 4, 2)
    bad = tot = 0
    for r0 in (0, 5000, H - 8):
        with torch.no_grad():
            y = O.forward(p, torch.from_numpy(O.features(msb, D, ofl, r0, r0 + 8)))
        res = torch.round(y * (2 ** K - 1)).numpy().reshape(8, W, 4).transpose(2, 0, 1)
        ref = np.round((msb[:, r0:r0 + 8].astype(np.uint16) << K).astype(np.float32) + res).astype(np.uint16)
        got = whole[:, r0:r0 + 8].view(torch.int16).cpu().numpy().view(np.uint16)
        diff = np.abs(got.astype(np.int64) - ref.astype(np.int64))
        assert diff.max() <= 1
        bad += int((diff != 0).sum())
        tot += diff.size
    assert bad <= max(1, int(1e-4 * tot)), (bad, tot)
