"""Reader for TensorFlow-1.x "bundle" checkpoints (`*.ckpt.index` + `*.ckpt.data-00000-of-00001`) without
TensorFlow -- enough to load the reference's `pretrained_model/pretrained_model.ckpt` into the flat
parameter dict of params.py.

Format: the `.index` file is an uncompressed LevelDB table -- a 48-byte footer (two varint BlockHandles +
magic) pointing at an index block whose entries point at data blocks; blocks hold prefix-compressed
(shared, non_shared, value_len) entries followed by a restart array and a 1-byte type + 4-byte crc trailer.
Keys are variable names, values are `BundleEntryProto`s (field 1 dtype, 2 shape, 3 shard_id, 4 offset,
5 size) locating the raw little-endian tensor bytes in the `.data` file.
"""
import os
import struct

import numpy as np
import torch

_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 9: np.int64}
_MAGIC = 0xdb4775248b80fb57


def _varint(buf, pos):
    result, shift = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7


def _block_entries(block):
    """(key, value) pairs of one table block (restart array ignored: entries are walked sequentially)."""
    num_restarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    limit = len(block) - 4 - 4 * num_restarts
    pos, key = 0, b""
    while pos < limit:
        shared, pos = _varint(block, pos)
        non_shared, pos = _varint(block, pos)
        vlen, pos = _varint(block, pos)
        key = key[:shared] + bytes(block[pos:pos + non_shared])
        pos += non_shared
        yield key, bytes(block[pos:pos + vlen])
        pos += vlen


def _read_block(data, offset, size):
    if data[offset + size] != 0:
        raise ValueError("compressed LevelDB block (type %d) is not supported" % data[offset + size])
    return data[offset:offset + size]


def _parse_entry(value):
    """BundleEntryProto -> (dtype enum, shape, offset, size)."""
    pos, dtype, shape, offset, size = 0, 0, [], 0, 0
    while pos < len(value):
        tag, pos = _varint(value, pos)
        field, wire = tag >> 3, tag & 7
        if wire == 0:
            v, pos = _varint(value, pos)
            if field == 1:
                dtype = v
            elif field == 4:
                offset = v
            elif field == 5:
                size = v
        elif wire == 2:
            n, pos = _varint(value, pos)
            sub = value[pos:pos + n]
            pos += n
            if field == 2:                       # TensorShapeProto { repeated Dim dim = 2 { int64 size = 1 } }
                sp = 0
                while sp < len(sub):
                    t2, sp = _varint(sub, sp)
                    if t2 & 7 == 2:
                        n2, sp = _varint(sub, sp)
                        dim, dp = sub[sp:sp + n2], 0
                        sp += n2
                        while dp < len(dim):
                            t3, dp = _varint(dim, dp)
                            if t3 & 7 == 0:
                                v3, dp = _varint(dim, dp)
                                if t3 >> 3 == 1:
                                    shape.append(v3)
                            else:
                                n3, dp = _varint(dim, dp)
                                dp += n3
                    else:
                        _, sp = _varint(sub, sp)
        elif wire == 5:
            pos += 4
        elif wire == 1:
            pos += 8
        else:
            raise ValueError("unexpected wire type %d" % wire)
    return dtype, shape, offset, size


def read_bundle(prefix):
    """All tensors of a TF-1.x checkpoint as {variable name: numpy array}."""
    index = open(prefix + ".index", "rb").read()
    if struct.unpack_from("<Q", index, len(index) - 8)[0] != _MAGIC:
        raise ValueError("%s.index is not a LevelDB table" % prefix)
    footer = index[-48:]
    pos = 0
    _, pos = _varint(footer, pos)                # metaindex handle
    _, pos = _varint(footer, pos)
    idx_off, pos = _varint(footer, pos)
    idx_size, pos = _varint(footer, pos)
    data_path = prefix + ".data-00000-of-00001"
    raw = np.memmap(data_path, dtype=np.uint8, mode="r")
    out = {}
    for _, handle in _block_entries(_read_block(index, idx_off, idx_size)):
        off, p = _varint(handle, 0)
        size, _ = _varint(handle, p)
        for key, value in _block_entries(_read_block(index, off, size)):
            if not key:                          # the header entry (BundleHeaderProto)
                continue
            dtype, shape, offset, nbytes = _parse_entry(value)
            if dtype not in _DTYPES:
                continue
            arr = np.frombuffer(raw[offset:offset + nbytes].tobytes(), dtype=_DTYPES[dtype]).reshape(shape)
            out[key.decode("utf-8")] = arr
    return out


def resolve_prefix(path):
    """Checkpoint prefix (what read_bundle opens as prefix + '.index') from a prefix or a checkpoint directory:
    a directory is resolved through its `checkpoint` file (model_checkpoint_path) or its only *.index file."""
    import glob
    import os
    import re
    if os.path.exists(path + ".index"):
        return path
    if os.path.isdir(path):
        state = os.path.join(path, "checkpoint")
        if os.path.exists(state):
            m = re.search(r'^model_checkpoint_path:\s*"([^"]+)"', open(state).read(), re.M)
            if m:
                cand = os.path.join(path, os.path.basename(m.group(1)))
                if os.path.exists(cand + ".index"):
                    return cand
        found = sorted(glob.glob(os.path.join(path, "*.index")))
        if len(found) == 1:
            return found[0][:-len(".index")]
        raise FileNotFoundError("%s: %d *.index files and no usable `checkpoint` file" % (path, len(found)))
    raise FileNotFoundError("%s: neither a checkpoint prefix (%s.index) nor a directory" % (path, path))


def load_reference_checkpoint(prefix, dtype=torch.float32):
    """The reference's checkpoint as the flat parameter dict of params.py: optimizer slots dropped,
    conv kernels squeezed from [1,1,Cin,Cout] / [1,Cin,Cout] to (Cin,Cout).  Also returns the global step."""
    tensors = read_bundle(prefix)
    P, step = {}, None
    for name, arr in tensors.items():
        if "/Adam" in name or name.startswith("beta1_power") or name.startswith("beta2_power"):
            continue
        if name == "Variable":
            step = int(arr)
            continue
        t = torch.from_numpy(np.array(arr)).to(dtype)
        if name.endswith("/weights"):
            t = t.reshape(t.shape[-2], t.shape[-1])
        P[name] = t
    return P, step
