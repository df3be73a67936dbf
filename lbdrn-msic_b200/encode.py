"""LBDRN encoder CLI -- drop-in for the reference's encode.py: same flags (-K -D -bc -nl -lr -bs -e -sr -prec -i -o
-vd --seed -rn), output-directory naming, log lines and `.bin` layout (reference encode.py:167-284).

The optimisation loop (reference encode.py:67-117 + modified_ignite_engine.py) runs as fused persistent CUDA kernels
(lbdrn_fused.FusedTrainer): the scene stays on the GPU as integer planes, one kernel launch per epoch does gather,
forward, MSE, backward and Adam for every batch, one launch per epoch evaluates the full-scene MSE, and the
best-epoch parameters are snapshotted on the device.  fpzip weight coding, GDAL I/O and the JPEG-2000 base layer stay on
the host, unchanged.
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
from lbdrn_container import write_image_header
from LBDRNdataset import LBDRNDataset, split_image
from LBDRNmodel import LBDRNModel


def sh(cmd, input=''):
    r = subprocess.run(cmd, shell=True, stdout=subprocess.PIPE, stderr=subprocess.PIPE, input=input.encode('utf-8'))
    assert r.returncode == 0, r.stderr.decode('utf-8')
    return r.stdout.decode('utf-8')


def write_scalars(log_dir, filename, result, args):
    """The reference's TensorBoard event file (encode.py:89,95,107): `train/loss/<name>` per iteration and `val/MSE/<name>`
    per evaluated epoch, same tags and steps.  The reference writes the loss scalar inside the loop, which forces a device
    synchronisation per step (encode.py:95); here the per-step losses come back as one vector per epoch and are flushed
    after the run.  Skipped silently when tensorboard is not installed (the reference lists it in requirements.txt)."""
    try:
        from torch.utils.tensorboard import SummaryWriter
    except Exception:                                                  # noqa: BLE001 -- optional dependency
        return False
    writer = SummaryWriter(log_dir=log_dir)
    for it, loss in enumerate(result['losses'], start=1):
        writer.add_scalar(f'train/loss/{filename}', loss, it)
    step = min(args.val_duration, args.epochs)
    for i, mse in enumerate(result['val_mse'], start=1):
        writer.add_scalar(f'val/MSE/{filename}', mse, i * step)
    writer.close()
    return True


def train(args):
    """Overfit one network to the raster at args.path; leaves `<name>_nn.bin` and `<name>_base.jp2` in output_dir."""
    import lbdrn_fused
    name = os.path.splitext(os.path.basename(args.path))[0]
    dataset = LBDRNDataset(args)                                       # writes <name>_base.tif
    model = LBDRNModel(dim_in=dataset.n_feature, dim_hidden=args.base_channel, dim_out=dataset.channels,
                       num_layers=args.num_layers)
    logger.log.info(model)
    for key, val in model.state_dict().items():
        logger.log.info('{}\t {}'.format(key, val.size()))
    logger.log.info('total_params: {}'.format(sum(p.numel() for p in model.parameters())))

    def report(epoch, mse, improved):
        if improved:
            logger.log.info('Save current best val model (MSE: {:.5f}) @epoch {}'.format(mse, epoch))
        else:
            logger.log.info('Model is not updated (MSE: {:.5f}) @epoch: {}'.format(mse, epoch))

    trainer = lbdrn_fused.FusedTrainer(model, dataset.scene, args.D, args.lr, args.batch_size, args.epochs,
                                       val_duration=args.val_duration, flags=dataset.flags,
                                       sampler=getattr(args, 'sampler', 'reference'), on_epoch=report)
    result = trainer.run()
    trainer.close()
    logger.log.info('best epoch: {}'.format(result['best_epoch']))
    write_scalars(args.output_dir, name, result, args)

    params = result['params'].numpy().reshape(-1)                      # state_dict order, C order
    nn_path = f'{args.output_dir}/{name}_nn.bin'
    with open(nn_path, 'wb') as f:
        f.write(fpzip.compress(params, precision=args.precision, order='C'))
    nn_bytes = os.path.getsize(nn_path)
    logger.log.info(f'nn: {nn_bytes} bytes, bpsp={nn_bytes * 8 / dataset.n_subpixels}')

    base_path, jp2_path = f'{args.output_dir}/{name}_base.tif', f'{args.output_dir}/{name}_base.jp2'
    logger.log.info(sh(f"gdal_translate -of JP2OpenJPEG -co QUALITY=100 -co REVERSIBLE=YES {base_path} {jp2_path}"))
    os.remove(base_path)
    base_bytes = os.path.getsize(jp2_path)
    logger.log.info(f"MSB: {base_bytes} bytes: bpsp={base_bytes * 8 / dataset.n_subpixels}")
    return result


def build_parser():
    p = argparse.ArgumentParser(description='LBDRN-MSIC')
    p.add_argument('--seed', type=int, default=19920517)
    p.add_argument('-rn', '--randomness', action='store_true', help='Allow randomness during training?')
    p.add_argument('-i', '--path', type=str, help='path of input tif or img file')
    p.add_argument('-o', '--output_dir', default='outputs', type=str, help='output dir')
    p.add_argument('-sr', '--split_ratio', type=int, default=1, help='tile size (default: 1)')
    p.add_argument('-K', '--K', type=int, default=5, help=' (default: 5)')
    p.add_argument('-bc', '--base_channel', type=int, default=64, choices=[32, 64, 128, 256],
                   help='base channel (default: 64); the fused kernels are built for 32 / 64 / 128 / 256')
    p.add_argument('-nl', '--num_layers', type=int, default=2, help='Number of layers (default: 2)')
    p.add_argument('-D', '--D', type=int, default=2, help='#neighbors (2D+1)^2')
    p.add_argument('-prec', '--precision', type=int, default=16, help=' (default: 16)')
    p.add_argument('-lr', '--lr', type=float, default=1e-3, help='learning rate (default: 1e-3)')
    p.add_argument('-bs', '--batch_size', type=int, default=8192, help='batch size (default: 8192)')
    p.add_argument('-e', '--epochs', type=int, default=10, help='number of epochs to train (default: 10)')
    p.add_argument('-vd', '--val_duration', type=int, default=1, help='number of epoch duration for val (default: 1)')
    p.add_argument('--sampler', choices=['reference', 'device'], default='reference',
                   help="batch order: 'reference' reproduces the reference DataLoader's permutations from the seed; "
                        "'device' draws them on the GPU (added flag)")
    return p


def main(argv=None):
    args = build_parser().parse_args(argv)
    if not args.randomness:
        torch.manual_seed(args.seed)
        torch.backends.cudnn.deterministic, torch.backends.cudnn.benchmark = True, False
        np.random.seed(args.seed)
        random.seed(args.seed)

    org_path = args.path
    name = os.path.splitext(os.path.basename(org_path))[0]
    args.output_dir = '{}/{}_r{}_K{}_bc{}_nl{}_D{}_prec{}_lr{}_bs{}_e{}'.format(
        args.output_dir, name, args.split_ratio, args.K, args.base_channel, args.num_layers, args.D, args.precision,
        args.lr, args.batch_size, args.epochs)
    os.makedirs(args.output_dir, exist_ok=True)
    bitstream_path = f'{args.output_dir}/{name}.bin'
    log_path = f'{args.output_dir}/encode.txt'
    if os.path.exists(log_path) and os.path.exists(bitstream_path):
        with open(log_path) as f:
            if "Time elapsed" in f.read():
                print('Bitstream already created!')
                sys.exit()
    logger.create_logger(args.output_dir, 'encode.txt')
    start_time = time.time()
    src = gdal.Open(org_path)
    width, height = src.RasterXSize, src.RasterYSize
    src = None

    stems = []
    if args.split_ratio > 1:
        split_image(args.path, args.output_dir, args.split_ratio)
        for i in range(args.split_ratio):
            for j in range(args.split_ratio):
                args.path = f'{args.output_dir}/tile_{i}_{j}.tif'
                logger.log.info(args)
                train(args)
                os.remove(args.path)
                stems.append(f'tile_{i}_{j}')
    else:
        logger.log.info(args)
        train(args)
        stems.append(name)

    nn_paths = [f'{args.output_dir}/{s}_nn.bin' for s in stems]
    base_paths = [f'{args.output_dir}/{s}_base.jp2' for s in stems]
    header_path = f'{bitstream_path}_header'
    write_image_header(header_path, args.split_ratio, width, height, args.K, args.base_channel, args.num_layers,
                       args.D, [os.path.getsize(p) for p in nn_paths], [os.path.getsize(p) for p in base_paths])
    with open(bitstream_path, 'wb') as out:                            # header, then nn_ij || base_ij per tile
        for p in [header_path] + [q for pair in zip(nn_paths, base_paths) for q in pair]:
            with open(p, 'rb') as f:
                out.write(f.read())
            os.remove(p)
    logger.log.info(f'Time elapsed: {time.time() - start_time}')


if __name__ == '__main__':
    main()
