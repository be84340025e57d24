"""Container / z-index wire format, byte-identical to the reference
(/root/reference/src/modules/entropy/utils.py:7-16, 95-132 and
 models/sd15_onedc_codec_stage1/codec_module.py:403-409, 426-429).

Container: big-endian u32 H, W, len(y stream), len(caption); y stream; z stream; caption.
z stream : hz*wz indices of `unit` (=14) bits each, MSB first, left-padded with zero bits to whole bytes.
"""
import math
import struct

import numpy as np


def get_padding_size(height, width, p=64):
    new_h = (height + p - 1) // p * p
    new_w = (width + p - 1) // p * p
    return 0, new_w - width, 0, new_h - height            # left, right, top, bottom


def encode_i(pic_height, pic_width, bit_stream_y, bit_stream_z, bit_stream_caption=b"", caption_length=0):
    if isinstance(bit_stream_caption, str):
        bit_stream_caption = bit_stream_caption.encode("utf-8")
    return (struct.pack(">4I", pic_height, pic_width, len(bit_stream_y), caption_length)
            + bytes(bit_stream_y) + bytes(bit_stream_z) + bytes(bit_stream_caption))


def decode_i(data, index_unit_length=14, ds=64):
    if len(data) < 16:
        raise ValueError("stream shorter than the 16-byte header")
    height, width, len_y, len_cap = struct.unpack(">4I", data[:16])
    pl, pr, pt, pb = get_padding_size(height, width, ds)
    pad_h, pad_w = height + pt + pb, width + pl + pr
    len_z = math.ceil((pad_h // ds) * (pad_w // ds) * index_unit_length / 8.0)
    if len(data) < 16 + len_y + len_z + len_cap:
        raise ValueError("truncated stream")
    o = 16
    y = data[o:o + len_y]
    o += len_y
    z = data[o:o + len_z]
    o += len_z
    return {"height": height, "width": width, "pad_height": pad_h, "pad_width": pad_w,
            "pad_tuple": (pl, pr, pt, pb), "bit_stream_y": y, "bit_stream_z": z,
            "bit_stream_caption": data[o:o + len_cap]}


def pack_indices(idx, unit=14):
    idx = np.asarray(idx, dtype=np.int64).reshape(-1)
    bits = ((idx[:, None] >> np.arange(unit - 1, -1, -1)) & 1).astype(np.uint8).reshape(-1)
    pad = (-len(bits)) % 8
    return np.packbits(np.concatenate([np.zeros(pad, np.uint8), bits])).tobytes()


def unpack_indices(data, count, unit=14):
    bits = np.unpackbits(np.frombuffer(data, dtype=np.uint8))
    need = count * unit
    if len(bits) < need:                                   # int.from_bytes semantics: missing high bits are zero
        bits = np.concatenate([np.zeros(need - len(bits), np.uint8), bits])
    bits = bits[len(bits) - need:].reshape(count, unit).astype(np.int64)
    return (bits << np.arange(unit - 1, -1, -1)).sum(axis=1).astype(np.int32)
