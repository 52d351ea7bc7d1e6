"""Test helper: numpy restatement of the packed-value CSC layout."""
import numpy as np


def pack_reference(W):
    """numpy restatement of the packed-value CSC layout: merged rows ascending per
    column, five base-3 digits d = v+1 per byte, little-endian, pad digit 1."""
    K, N = W.shape
    col_ptr = np.zeros(N + 1, np.int32)
    rows, digs = [], []
    for n in range(N):
        k = np.flatnonzero(W[:, n])
        rows.append(k.astype(np.int32))
        digs.append((W[k, n] + 1).astype(np.int64))
        col_ptr[n + 1] = col_ptr[n] + k.size
    row_idx = np.concatenate(rows) if rows else np.zeros(0, np.int32)
    d = np.concatenate(digs) if digs else np.zeros(0, np.int64)
    pad = (-d.size) % 5
    d = np.concatenate([d, np.ones(pad, np.int64)]).reshape(-1, 5)
    vals = (d * np.array([1, 3, 9, 27, 81])).sum(axis=1).astype(np.uint8)
    return col_ptr, row_idx, vals


def pack_reference_csr(W):
    """Packed-value CSR = the packed CSC layout of the transpose: row_ptr[K+1], col_idx ascending
    per row, the same five-digits-per-byte packing along the merged entry list."""
    return pack_reference(np.ascontiguousarray(W.T))
