"""LBDRN decoder CLI -- drop-in for the reference's decode.py: same flags (-i, -org, --seed), log lines
(`MSE: `, `PSNR: `, `Total size: N bytes, bpsp=`, `Time elapsed: `) and `.bin` layout (reference decode.py:151-224).

Inference (reference decode.py:77-134: host feature matrix + chunked model calls + numpy reconstruction) is ONE fused
CUDA kernel over the base layer: tile+halo staging, features in registers/smem, the SIREN MLP, sigmoid, inverse
quantisation and the integer write.  fpzip and the JPEG-2000 base layer stay on the host, unchanged.
"""
import argparse
import os
import random
import subprocess
import sys
import time

try:
    import fpzip                       # the reference's dependency (requirements.txt:2): used whenever it is installed
except ImportError:                    # same algorithm from liblbdrn_b200 (csrc/lbdrn_fpz.cpp)
    import lbdrn_fpzip as fpzip
import numpy as np
import torch
from osgeo import gdal

import logger
from lbdrn_container import read_image_header
from LBDRNdataset import merge_tiles, write_tiff_with_gdal

gdal.UseExceptions()


def sh(cmd, input=''):
    r = subprocess.run(cmd, shell=True, stdout=subprocess.PIPE, stderr=subprocess.PIPE, input=input.encode('utf-8'))
    assert r.returncode == 0, r.stderr.decode('utf-8')
    return r.stdout.decode('utf-8')


K = D = bc = nl = None   # header fields; module globals like in the reference's decode.py (set by main)


def test(bitstream, dirname, filename, nn_bytes, base_bytes):
    """Decode one tile from the front of `bitstream` into `<dirname>/<filename>_recon.tif`; returns the rest."""
    import lbdrn_fused
    nn_stream, bitstream = bitstream[:nn_bytes], bitstream[nn_bytes:]
    base_stream, bitstream = bitstream[:base_bytes], bitstream[base_bytes:]
    recon_path, jp2_path = f'{dirname}/{filename}_recon.tif', f'{dirname}/{filename}_base.jp2'
    with open(jp2_path, 'wb') as f:
        f.write(base_stream)
    logger.log.info(sh(f"gdal_translate -of GTiff {jp2_path} {recon_path}"))
    base = gdal.Open(recon_path).ReadAsArray()
    base = base.reshape((-1,) + base.shape[-2:])                        # CHW; u8 stays u8 on the device
    params = np.asarray(fpzip.decompress(nn_stream, order='C')[0][0][0], dtype=np.float32)
    image = lbdrn_fused.decode_image(base, params, K, D, bc, nl)         # flags from constants.py, like the reference
    write_tiff_with_gdal(recon_path, image)
    logger.log.info(f'Recon: {recon_path}')
    for p in (jp2_path, jp2_path + '.aux.xml'):
        if os.path.exists(p):
            os.remove(p)
    return bitstream


def main(argv=None):
    p = argparse.ArgumentParser(description='LBDRN-RSIC')
    p.add_argument('--seed', type=int, default=19920517)
    p.add_argument('-i', '--bin_path', type=str, help='binstream path')
    p.add_argument('-org', '--org_path', type=str, default=None, help='org path')
    args = p.parse_args(argv)
    torch.manual_seed(args.seed)
    np.random.seed(args.seed)
    random.seed(args.seed)

    dirname, basename = os.path.split(args.bin_path)
    filename = os.path.splitext(basename)[0]
    if os.path.exists(f'{dirname}/decode.txt'):
        with open(f'{dirname}/decode.txt') as f:
            if "bpsp" in f.read():
                print('Bitstream already decoded!')
                sys.exit()
    logger.create_logger(dirname, 'decode.txt')
    logger.log.info(f'Binstream: {args.bin_path}')
    start_time = time.time()
    with open(args.bin_path, 'rb') as f:
        bitstream = f.read()
    global K, D, bc, nl
    n_hdr, split_ratio, width, height, K, bc, nl, D, nn_list, base_list = read_image_header(bitstream)
    bitstream = bitstream[n_hdr:]
    recon_path = f'{dirname}/{basename[:-4]}_recon.tif'
    if split_ratio > 1:
        for i in range(split_ratio):
            for j in range(split_ratio):
                t = i * split_ratio + j
                bitstream = test(bitstream, dirname, f'tile_{i}_{j}', nn_list[t], base_list[t])
        merge_tiles(dirname, recon_path, split_ratio, width, height)
        for i in range(split_ratio):
            for j in range(split_ratio):
                os.remove(f'{dirname}/tile_{i}_{j}_recon.tif')
    else:
        bitstream = test(bitstream, dirname, filename, nn_list[0], base_list[0])
    logger.log.info(f'Time elapsed: {time.time() - start_time}')

    if args.org_path is not None:
        org = gdal.Open(args.org_path).ReadAsArray()
        rec = gdal.Open(recon_path).ReadAsArray()
        n_bytes = os.path.getsize(args.bin_path)
        if org.dtype == np.uint16 and rec.dtype == np.uint16 and torch.cuda.is_available():
            import lbdrn_fused
            mse = np.float32(lbdrn_fused.image_mse(org, rec))            # exact integer sum on the device
        else:
            mse = np.mean((org.astype(np.float32) - rec.astype(np.float32)) ** 2)
        logger.log.info(f"MSE: {mse}")
        logger.log.info(f"PSNR: {10 * np.log10(10000 ** 2 / mse)}")       # peak fixed at 10000 (decode.py:218)
        logger.log.info(f"Total size: {n_bytes} bytes, bpsp={n_bytes * 8 / np.prod(org.shape)}")
        os.remove(recon_path)


if __name__ == '__main__':
    main()
