"""The `.bin` container of the codec (reference encode.py:29-64, decode.py:25-53): byte-exact reader / writer.

    [1B header length][1B split_ratio][2B width][2B height][1B K<<4 | D][1B log2(bc)<<4 | nl]
    [3B x sr^2 nn sub-stream sizes][4B x sr^2 base sub-stream sizes]          (all big-endian, unsigned)
    then, per tile in row-major (i, j) order:  nn_ij || base_ij
"""
import struct


def pack_header(split_ratio, width, height, K, bc, nl, D, nn_bytes_list, base_bytes_list):
    log2bc = bc.bit_length() - 1
    if bc < 1 or (1 << log2bc) != bc:
        raise ValueError("base_channel must be a power of two (the header stores log2(bc) in 4 bits)")
    n = 8 + 3 * len(nn_bytes_list) + 4 * len(base_bytes_list)
    for v, lim, what in ((n, 255, "header length"), (split_ratio, 255, "split_ratio"), (K, 15, "K"), (D, 15, "D"),
                         (log2bc, 15, "log2(bc)"), (nl, 15, "nl"), (width, 65535, "width"), (height, 65535, "height")):
        if not 0 <= v <= lim:
            raise OverflowError(f"{what}={v} does not fit its header field")
    out = struct.pack(">BBHHBB", n, split_ratio, width, height, (K << 4) | D, (log2bc << 4) | nl)
    for v in nn_bytes_list:
        out += v.to_bytes(3, "big")
    for v in base_bytes_list:
        out += v.to_bytes(4, "big")
    return out


def write_image_header(header_path, split_ratio, width, height, K, bc, nl, D, nn_bytes_list, base_bytes_list):
    blob = pack_header(split_ratio, width, height, K, bc, nl, D, nn_bytes_list, base_bytes_list)
    with open(header_path, "wb") as f:
        f.write(blob)
    return len(blob)


def read_image_header(bitstream):
    """-> (n_bytes_header, split_ratio, width, height, K, bc, nl, D, nn_bytes_list, base_bytes_list)"""
    n, sr, width, height, kd, bcnl = struct.unpack(">BBHHBB", bitstream[:8])
    tiles = sr * sr
    nn = [int.from_bytes(bitstream[8 + 3 * t:11 + 3 * t], "big") for t in range(tiles)]
    off = 8 + 3 * tiles
    base = [int.from_bytes(bitstream[off + 4 * t:off + 4 * t + 4], "big") for t in range(tiles)]
    return n, sr, width, height, kd >> 4, 1 << (bcnl >> 4), bcnl & 15, kd & 15, nn, base
