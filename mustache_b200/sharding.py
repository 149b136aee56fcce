"""Multi-GPU sharding of the block loop (host logic; one process per GPU over torch.distributed).

The reference fans the blocks of ONE chromosome out to `-p` OS processes and walks the chromosomes serially
(mustache.py:913-937, 1057).  Blocks are independent (own borders, own exponential fits, own BH; mustache.py:697-850),
so here the blocks of ALL requested chromosomes form one pool that is split evenly over the ranks (every block costs the
same on the GPU: the kernels sweep the whole band whatever the contact count, so LPT degenerates to equal counts):

  1. every chromosome has ONE owner rank that reads and normalises it (nobody repeats that work);
  2. `balanced_assignment` gives every rank floor/ceil(total / world) blocks, owners keeping their own blocks first;
  3. `exchange_blocks` moves the block-local COO of the blocks an owner cannot keep to the rank that computes them
     (the path's one real exchange step: all_to_all_single, NCCL over NVLink on GPUs, gloo in the CPU tests);
  4. every rank runs its blocks through its engine and post-processes them (BH stays per block, mustache.py:774-779);
  5. `gather_loops` collects the surviving calls on the rank that writes the TSV (replaces Manager().list(),
     mustache.py:913-914, 959).
"""
import numpy as np


def chromosome_owners(n_chrom, world, sizes=None):
    """Owner rank per chromosome: round-robin, longest first when sizes are known (so that reading and normalising --
    host work proportional to the chromosome length -- is spread evenly too)."""
    order = list(range(n_chrom))
    if sizes is not None:
        order.sort(key=lambda c: (-int(sizes[c]), c))
    load = [0] * world
    owners = [0] * n_chrom
    for k, c in enumerate(order):
        r = min(range(world), key=lambda q: (load[q], q)) if sizes is not None else k % world
        owners[c] = r
        load[r] += int(sizes[c]) if sizes is not None else 1
    return owners


def balanced_assignment(blocks_per_chrom, owners, world):
    """{rank: [(chrom, block), ...]}: every rank gets floor or ceil(total / world) blocks; an owner keeps as many of its own
    blocks as its quota allows, the surplus goes to the ranks that still have room, in rank order.  Deterministic, so every
    rank computes the same table without communication."""
    total = int(sum(blocks_per_chrom))
    quota = [total // world + (1 if r < total % world else 0) for r in range(world)]
    out = {r: [] for r in range(world)}
    surplus = []
    for c, nb in enumerate(blocks_per_chrom):
        r = owners[c]
        for b in range(nb):
            if len(out[r]) < quota[r]:
                out[r].append((c, b))
            else:
                surplus.append((c, b))
    r = 0
    for item in surplus:
        while len(out[r]) >= quota[r]:
            r += 1
        out[r].append(item)
    return out


def _pack_blocks(items):
    """[(chrom, block, rows, cols, vals)] -> int64 words: per block a 3-word header (chrom, block, nnz), nnz words of
    (row << 32 | col), nnz words holding the bits of the float64 values."""
    parts = []
    for chrom, block, rows, cols, vals in items:
        m = len(vals)
        parts.append(np.array([chrom, block, m], dtype=np.int64))
        parts.append((np.asarray(rows, np.int64) << 32) | np.asarray(cols, np.int64))
        parts.append(np.ascontiguousarray(vals, dtype=np.float64).view(np.int64))
    return np.concatenate(parts) if parts else np.zeros(0, np.int64)


def _unpack_blocks(words):
    out, pos = [], 0
    while pos < len(words):
        chrom, block, m = (int(t) for t in words[pos:pos + 3])
        pos += 3
        rc = words[pos:pos + m]
        vals = words[pos + m:pos + 2 * m].view(np.float64)
        pos += 2 * m
        out.append((chrom, block, (rc >> 32).astype(np.int64), (rc & 0xFFFFFFFF).astype(np.int64), vals.copy()))
    return out


def exchange_blocks(send, rank, world, device):
    """send: {dst rank: [(chrom, block, rows, cols, vals)]} of the blocks this rank owns but another rank computes.
    Returns the list of blocks other ranks sent here.  Two all_to_all_single calls: word counts, then the payload."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return []
    chunks = [_pack_blocks(send.get(d, [])) for d in range(world)]
    in_split = [len(c) for c in chunks]
    cnt_in = torch.tensor(in_split, dtype=torch.int64, device=device)
    cnt_out = torch.zeros(world, dtype=torch.int64, device=device)
    dist.all_to_all_single(cnt_out, cnt_in)
    out_split = [int(t) for t in cnt_out.tolist()]
    payload = torch.from_numpy(np.concatenate(chunks) if sum(in_split) else np.zeros(0, np.int64)).to(device)
    recv = torch.zeros(sum(out_split), dtype=torch.int64, device=device)
    dist.all_to_all_single(recv, payload, output_split_sizes=out_split, input_split_sizes=in_split)
    return _unpack_blocks(recv.cpu().numpy())


def gather_loops(rows, rank, world, device, root=0):
    """rows: float64 array [m, w] of this rank's calls (chromosome index, x, y, fdr, scale[, tag]); the root receives the
    concatenation in rank order, the others an empty array.  Counts all_gather, then one gather of padded rows."""
    import torch
    import torch.distributed as dist
    rows = np.ascontiguousarray(rows, dtype=np.float64)
    if world == 1:
        return rows
    w = rows.shape[1]
    cnt = torch.tensor([rows.shape[0]], dtype=torch.int64, device=device)
    cnts = torch.zeros(world, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(cnts, cnt)
    sizes = [int(t) for t in cnts.tolist()]
    mx = max(max(sizes), 1)
    buf = torch.zeros((mx, w), dtype=torch.float64, device=device)
    if rows.shape[0]:
        buf[:rows.shape[0]] = torch.from_numpy(rows).to(device)
    outs = [torch.empty_like(buf) for _ in range(world)] if rank == root else None
    dist.gather(buf, outs, dst=root)
    if rank != root:
        return np.zeros((0, w))
    return np.concatenate([o[:s].cpu().numpy() for o, s in zip(outs, sizes)], axis=0)


def all_gather_meta(obj, world):
    """Small python objects (chromosome lengths) from every rank."""
    import torch.distributed as dist
    if world == 1:
        return [obj]
    out = [None] * world
    dist.all_gather_object(out, obj)
    return out
