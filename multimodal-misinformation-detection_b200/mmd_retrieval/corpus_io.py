"""Corpus containers -> prepared, HBM-resident corpora.

The reference keeps its corpora in two on-disk formats and re-normalises them at query time:

  text   h5 file with `embeddings` fp16 [N,768] and `ids` (b"train_17", ...)   src/evidence/text2text_retrieval.py:39-47,146-155
  image  pickle of dict[path -> FloatTensor[2048]] in insertion order         src/evidence/im2im_retrieval.py:51-62

Here a corpus is streamed chunk by chunk through K1 (normalise + cast) straight into its prepared operand tiles, so
the full-precision matrix never has to exist on the device (a 100M x 768 corpus is 307 GB in fp32 but 9.6 GB per GPU as
fp8 shards) and nothing is re-normalised per query.  File readers are thin: h5py is imported only if present.
"""
from __future__ import annotations

import pickle
from typing import Callable, Hashable, Iterable, List, Optional, Sequence, Tuple, Union

import torch

from . import _lib, ops


def prepare_streamed(chunks: Iterable, n_rows: int, dim: int, dtype: str = "bf16", metric: str = "cos",
                     eps: float = ops.DEFAULT_EPS, keep_source: Union[bool, torch.dtype] = False, device=None,
                     idx_offset: int = 0) -> ops.PreparedCorpus:
    """Build a PreparedCorpus from an iterable of [rows_i, dim] chunks (tensors / ndarrays, host or device) whose row
    counts add up to n_rows.  keep_source: False, True (keep the chunks' dtype) or a torch dtype to store the source in
    (e.g. torch.float16, what the reference stores) for the exact re-score."""
    if device is not None:
        dev = torch.device(device)
    else:
        dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else None
    if dev is None:
        raise _lib.MmdError("no CUDA device is available; the retrieval path has no CPU fallback")
    ops._require_cuda(dev)
    _, row_bytes = ops.prepared_layout(dtype, dim)
    rows = torch.empty((n_rows, row_bytes), dtype=torch.uint8, device=dev)
    inv = torch.empty((n_rows,), dtype=torch.float32, device=dev)
    source = None
    lib = _lib.load()
    at = 0
    for chunk in chunks:
        x = ops._as_rows(chunk, dev)
        if x.shape[1] != dim:
            raise ValueError(f"chunk has dim {x.shape[1]}, expected {dim}")
        n = x.shape[0]
        if at + n > n_rows:
            raise ValueError(f"chunks hold more than n_rows={n_rows} rows")
        if n:
            with torch.cuda.device(dev):
                rc = lib.mmd_normalize_cast(ops._ptr(x), ops._SRC_DTYPE[x.dtype], n, dim, x.stride(0), int(metric == "cos"), float(eps),
                                            ops._OP_DTYPE[dtype], _lib.SIDE_CORPUS, ops._ptr(rows[at:]), ops._ptr(inv[at:]),
                                            ops._stream_ptr(dev))
            _lib.check(rc, "mmd_normalize_cast")
            if keep_source is not False:
                sdt = x.dtype if keep_source is True else keep_source
                if source is None:
                    source = torch.empty((n_rows, dim), dtype=sdt, device=dev)
                source[at:at + n] = x
        at += n
    if at != n_rows:
        raise ValueError(f"chunks hold {at} rows, expected {n_rows}")
    return ops.PreparedCorpus(rows=rows, inv_norm=inv if metric == "cos" else None, source=source, n=n_rows, dim=dim, op=dtype,
                              metric=metric, eps=eps, idx_offset=idx_offset)


def _row_chunks(arr, chunk_rows: int):
    for lo in range(0, arr.shape[0], chunk_rows):
        yield arr[lo:lo + chunk_rows]


def load_text_corpus(path: str, dtype: str = "bf16", metric: str = "cos", chunk_rows: int = 262144, rows: Optional[Tuple[int, int]] = None,
                     keep_source: Union[bool, torch.dtype] = True, device=None) -> Tuple[ops.PreparedCorpus, List[str]]:
    """Text-evidence corpus file -> (PreparedCorpus, ids).  `.h5` in the reference's layout (datasets `embeddings`, `ids`;
    needs h5py), or `.npz` / `.npy` with the same arrays (`.npy`: embeddings only, memory-mapped).  rows=(lo, hi) loads
    one shard of a row-sharded corpus; its idx_offset is lo."""
    import numpy as np
    ids: List[str] = []
    if path.endswith((".h5", ".hdf5")):
        try:
            import h5py
        except ImportError as e:
            raise _lib.MmdError("reading .h5 corpora needs h5py (not installed); convert to .npz / .npy") from e
        with h5py.File(path, "r") as f:
            emb = f["embeddings"]
            lo, hi = rows if rows is not None else (0, emb.shape[0])
            ids = [i.decode() if isinstance(i, bytes) else str(i) for i in f["ids"][lo:hi]] if "ids" in f else []
            pc = prepare_streamed((np.asarray(emb[a:min(a + chunk_rows, hi)]) for a in range(lo, hi, chunk_rows)), hi - lo,
                                  emb.shape[1], dtype, metric, ops.DEFAULT_EPS, keep_source, device, idx_offset=lo)
        return pc, ids
    if path.endswith(".npz"):
        z = np.load(path, allow_pickle=False)
        emb = z["embeddings"]
        all_ids = z["ids"] if "ids" in z.files else None
    else:
        emb = np.load(path, mmap_mode="r")
        all_ids = None
    lo, hi = rows if rows is not None else (0, emb.shape[0])
    if all_ids is not None:
        ids = [i.decode() if isinstance(i, bytes) else str(i) for i in all_ids[lo:hi]]
    pc = prepare_streamed((np.ascontiguousarray(c) for c in _row_chunks(emb[lo:hi], chunk_rows)), hi - lo, emb.shape[1], dtype,
                          metric, ops.DEFAULT_EPS, keep_source, device, idx_offset=lo)
    return pc, ids


def load_image_corpus(path: str, dtype: str = "bf16", chunk_rows: int = 65536, keep_source: Union[bool, torch.dtype] = True,
                      device=None) -> Tuple[ops.PreparedCorpus, List[Hashable]]:
    """The reference's image-feature pickle (dict[key -> 1-D tensor], insertion order = corpus row order) ->
    (PreparedCorpus with the image path's eps = 1e-6, keys)."""
    with open(path, "rb") as f:
        feature_dict = pickle.load(f)
    keys = list(feature_dict.keys())
    vals = list(feature_dict.values())
    if not keys:
        raise ValueError("empty feature dict")
    dim = int(torch.as_tensor(vals[0]).numel())

    def gen():
        for lo in range(0, len(vals), chunk_rows):
            yield torch.stack([torch.as_tensor(v).reshape(-1).float() for v in vals[lo:lo + chunk_rows]])

    pc = prepare_streamed(gen(), len(keys), dim, dtype, "cos", 1e-6, keep_source, device)
    return pc, keys
