"""The single collective of the path: variable-length gather of candidate records before BH-FDR.

Replaces the `multiprocessing.Manager().list()` result sink of the reference (mustache.py:913-914, 959): every rank
contributes the records of the blocks it processed; BH stays grouped per block (mustache.py:774-779), so records
carry (chromosome index, block index).  Two all_gathers over torch.distributed (NCCL on GPUs, gloo in CPU tests):
fixed-size counts, then records padded to the longest rank.  Payload is a few KB..MB; latency is the only cost.
"""
import numpy as np

FIELDS = ("chrom", "block", "row", "col", "v", "score_id", "p", "nz_count")


def pack_records(recs, chrom=0, block_ids=None, with_pair=False):
    """List of engine record dicts (one per block) -> float64 array [n, 8 (+1 with pPair)] (integers are exact in float64)."""
    parts = []
    width = len(FIELDS) + (1 if with_pair else 0)
    for k, r in enumerate(recs):
        b = k if block_ids is None else block_ids[k]
        m = len(r["rows"])
        a = np.empty((m, width), dtype=np.float64)
        if with_pair:
            a[:, len(FIELDS)] = r["pair"]
        a[:, 0] = chrom
        a[:, 1] = b
        a[:, 2] = r["rows"]
        a[:, 3] = r["cols"]
        a[:, 4] = r["v"]
        a[:, 5] = r["score_id"]
        a[:, 6] = r["p"]
        a[:, 7] = r["nz_count"]
        parts.append(a)
    return np.concatenate(parts, axis=0) if parts else np.zeros((0, width))


def all_gather_packed(packed, rank, world, device):
    """Every rank receives the concatenation of all ranks' packed records (rank order)."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return packed
    cnt = torch.tensor([packed.shape[0]], dtype=torch.int64, device=device)
    cnts = [torch.zeros_like(cnt) for _ in range(world)]
    dist.all_gather(cnts, cnt)
    sizes = [int(c.item()) for c in cnts]
    mx = max(max(sizes), 1)
    buf = torch.zeros((mx, packed.shape[1]), dtype=torch.float64, device=device)
    if packed.shape[0]:
        buf[:packed.shape[0]] = torch.from_numpy(packed).to(device)
    outs = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(outs, buf)
    return np.concatenate([o[:s].cpu().numpy() for o, s in zip(outs, sizes)], axis=0)


def all_gather_records(recs, rank, world, device, chrom=0, block_ids=None, with_pair=False):
    return all_gather_packed(pack_records(recs, chrom, block_ids, with_pair), rank, world, device)


def split_by_block(gathered):
    """{(chrom, block): record dict} from a gathered array."""
    out = {}
    if gathered.shape[0] == 0:
        return out
    keys = gathered[:, 0].astype(np.int64) * (1 << 32) + gathered[:, 1].astype(np.int64)
    order = np.lexsort((gathered[:, 3], gathered[:, 2], keys))
    g = gathered[order]
    keys = keys[order]
    bounds = np.flatnonzero(np.diff(keys)) + 1
    for seg in np.split(np.arange(len(keys)), bounds):
        a = g[seg]
        rec = dict(rows=a[:, 2].astype(np.int32), cols=a[:, 3].astype(np.int32), v=a[:, 4],
                   score_id=a[:, 5].astype(np.int32), p=a[:, 6], nz_count=int(a[0, 7]), n_found=len(seg))
        if a.shape[1] > len(FIELDS):
            rec["pair"] = a[:, len(FIELDS)]
        out[(int(a[0, 0]), int(a[0, 1]))] = rec
    return out


def all_gather_device(dev_recs, world):
    """NCCL gather of one block's records that never leaves the devices: `dev_recs` is ScaleSpaceEngine.records_device()
    (torch tensors aliasing the engine's buffers).  Counts first, then every field padded to the longest rank.
    Returns (list of per-rank counts, dict of gathered tensors [world, max_n])."""
    import torch
    import torch.distributed as dist
    dev = dev_recs["v"].device
    cnt = torch.tensor([dev_recs["n_found"], dev_recs["nz_count"]], dtype=torch.int64, device=dev)
    cnts = torch.empty(world * 2, dtype=torch.int64, device=dev)        # flat: gloo only takes 1-D outputs here
    dist.all_gather_into_tensor(cnts, cnt)
    sizes = cnts.view(world, 2)[:, 0].tolist()
    mx = max(max(sizes), 1)
    out = {}
    for name in ("rows", "cols", "v", "scored_index", "p"):
        src = dev_recs[name]
        pad = torch.zeros(mx, dtype=src.dtype, device=dev)
        pad[:src.numel()] = src
        buf = torch.empty(world * mx, dtype=src.dtype, device=dev)
        dist.all_gather_into_tensor(buf, pad)
        out[name] = buf.view(world, mx)
    return sizes, out


_DEV_FIELDS = ("rows", "cols", "v", "scored_index", "p")


def gather_device_to_root(dev_recs, world, rank, root=0):
    """The candidate gather as the path needs it: only the rank that runs BH-FDR receives.  One collective of counts,
    then ONE dist.gather of a byte buffer holding all five fields of the block (each padded to the longest rank), so the
    cost is one launch plus the root's ingest instead of five all_gathers that every rank pays for.
    Returns (sizes, {field: tensor [world, max_n]}) on the root, (sizes, None) elsewhere."""
    import torch
    import torch.distributed as dist
    dev = dev_recs["v"].device
    cnt = torch.tensor([dev_recs["n_found"], dev_recs["nz_count"]], dtype=torch.int64, device=dev)
    cnts = torch.empty(world * 2, dtype=torch.int64, device=dev)        # flat: gloo only takes 1-D outputs here
    dist.all_gather_into_tensor(cnts, cnt)
    sizes = cnts.view(world, 2)[:, 0].tolist()
    mx = max(max(sizes), 1)
    widths = [dev_recs[name].element_size() for name in _DEV_FIELDS]
    buf = torch.zeros(mx * sum(widths), dtype=torch.uint8, device=dev)
    off = 0
    for name, w in zip(_DEV_FIELDS, widths):
        src = dev_recs[name]
        buf[off:off + src.numel() * w] = src.contiguous().view(torch.uint8)
        off += mx * w
    outs = [torch.empty_like(buf) for _ in range(world)] if rank == root else None
    dist.gather(buf, outs, dst=root)
    if rank != root:
        return sizes, None
    stacked = torch.stack(outs)                                  # [world, mx * 28]
    fields, off = {}, 0
    for name, w in zip(_DEV_FIELDS, widths):
        fields[name] = stacked[:, off:off + mx * w].contiguous().view(dev_recs[name].dtype).view(world, mx)
        off += mx * w
    return sizes, fields
