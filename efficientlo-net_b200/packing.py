"""Host-side packing of folded weights into the layouts the kernels stream (csrc/elo_mlp.cuh)."""
import torch

from .params import fold_bn

CHUNK_FLOATS = 2048


def folded_chain(P, scopes, pad_hidden=None):
    """[(W', b')] of a layer chain, BN folded.  pad_hidden = 64: hidden layers narrower than that are widened with
    zero columns (zero bias: ReLU keeps them at 0) and the next layer gets matching zero rows -- the chain computes
    the same function on an engine that only knows 64- / 128-wide layers (the 32-32-64 set-conv of pyramid layer 2)."""
    out, extra_rows = [], 0
    for i, scope in enumerate(scopes):
        w, b = fold_bn(P, scope)
        w, b = w.float(), b.float()
        if extra_rows:
            w = torch.cat([w, torch.zeros(extra_rows, w.shape[1])], 0)
        extra_rows = 0
        if pad_hidden and i + 1 < len(scopes) and w.shape[1] < pad_hidden:
            extra_rows = pad_hidden - w.shape[1]
            w = torch.cat([w, torch.zeros(w.shape[0], extra_rows)], 1)
            b = torch.cat([b, torch.zeros(extra_rows)])
        out.append((w, b))
    return out


def pack_stream(P, scopes, pad_hidden=None):
    """Chunked stream for the shared-memory GEMM engine: per layer, rows 0..Cin-1 = W'[k][:], row Cin =
    folded bias, zero rows up to a multiple of 2048 / Cout; layers back to back in execution order."""
    parts = []
    for scope, (w, b) in zip(scopes, folded_chain(P, scopes, pad_hidden)):
        cin, cout = w.shape
        if cout not in (64, 128):
            raise ValueError("%s: the GEMM engine takes 64- or 128-wide layers, got %d" % (scope, cout))
        rows_per_chunk = CHUNK_FLOATS // cout
        rows = cin + 1
        padded = (rows + rows_per_chunk - 1) // rows_per_chunk * rows_per_chunk
        block = torch.zeros(padded, cout, dtype=torch.float32)
        block[:cin] = w.float()
        block[cin] = b.float()
        parts.append(block.reshape(-1))
    return torch.cat(parts).contiguous()


def split_tf32(w):
    """w = hi + lo with hi = w truncated to tf32 (low 13 mantissa bits cleared), lo = w - hi (exact in fp32)."""
    w = w.float().contiguous()
    hi = (w.view(torch.int32) & -8192).view(torch.float32)
    return hi, (w - hi)


def pack_stream_tc(P, scopes, pad_hidden=None):
    """Stream for the tensor-core engine (csrc/elo_tc_engine.cuh): per layer ceil(Cin / R) chunks of
    R = 2048 / Cout k-rows; a chunk is [hi | lo], each half in the canonical K-major core-matrix order
    [R/4][Cout][4] (k-rows beyond Cin are zero); after all chunks, the folded biases of all layers."""
    chunks, biases = [], []
    for scope, (w, b) in zip(scopes, folded_chain(P, scopes, pad_hidden)):
        cin, cout = w.shape
        if cout not in (64, 128):
            raise ValueError("%s: the GEMM engine takes 64- or 128-wide layers, got %d" % (scope, cout))
        R = CHUNK_FLOATS // cout
        nch = (cin + R - 1) // R
        padded = torch.zeros(nch * R, cout, dtype=torch.float32)
        padded[:cin] = w.float()
        hi, lo = split_tf32(padded)
        for c in range(nch):
            for part in (hi, lo):
                blk = part[c * R:(c + 1) * R].reshape(R // 4, 4, cout).permute(0, 2, 1).contiguous()
                chunks.append(blk.reshape(-1))
        biases.append(b.float().reshape(-1))
    return torch.cat(chunks + biases).contiguous()


def pack_plain(P, scopes):
    """Plain packing for the register-MLP set-conv: W1, b1, W2, b2, W3, b3 (row-major [Cin][Cout])."""
    parts = []
    for scope in scopes:
        w, b = fold_bn(P, scope)
        parts += [w.float().reshape(-1), b.float().reshape(-1)]
    return torch.cat(parts).contiguous()
