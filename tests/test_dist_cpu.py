"""world_size-2 gloo test (CPU) of the data-parallel host logic: sharded token statistics,
packed and all-reduced exactly like `_engine.statistics` packs them, give the same MP ranks and
mixing weights as one process on the concatenated batch.  Compute here is the CPU model (the
CUDA kernels need a GPU) -- what is under test is the sharding contract: additive statistics,
global row count, one flat buffer, shards that are slices of the global batch."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import basd_b200.synthetic as syn
from oracle import kernel_model as km
from tests import _cases as cs


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _pack(stats_s, stats_t):
    parts = [torch.stack([g for g, _ in stats_s]).flatten(), torch.stack([c for _, c in stats_s]).flatten(),
             torch.stack([g for g, _ in stats_t]).flatten(), torch.stack([c for _, c in stats_t]).flatten()]
    return torch.cat(parts), [p.numel() for p in parts]


def _unpack(flat, sizes, e, l, d_s, d_t):
    a, b, c, d = torch.split(flat, sizes)
    return (list(zip(a.view(e, d_s, d_s), b.view(e, d_s))), list(zip(c.view(l, d_t, d_t), d.view(l, d_t))))


def _worker(rank, world, port, local_batch, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from basd_b200 import _autograd, _engine
        assert _autograd.world_size(None) == world
        work = cs.workload("c1", local_batch)
        _, _, st, te, _ = syn.make_inputs(work, seed=4, batch_offset=rank * local_batch)
        layers = sorted(st)
        # the contract of _engine.statistics: every rank accumulates in its OWN frame (the rough mean of its
        # shard, bf16-rounded), ONE all-reduce sums the Grams and carries each rank's column sums and shift in a
        # slot of its own, then the statistics move to the common frame (center.cu: merge_shifted_stats_kernel)
        tensors = [st[l] for l in layers] + [te[k] for k in sorted(te)]
        shifts = [t.reshape(-1, t.shape[-1])[:256].mean(0).bfloat16().float() for t in tensors]
        shifted = [t - m for t, m in zip(tensors, shifts)]
        flat, sizes = _pack([km.token_stats(x) for x in shifted[:len(layers)]],
                            [km.token_stats(x) for x in shifted[len(layers):]])
        cols = torch.cat([km.token_stats(x)[1] for x in shifted])
        mus = torch.cat(shifts)
        slots = torch.zeros(2, world, cols.numel())
        slots[0, rank], slots[1, rank] = cols, mus
        buf = torch.cat([flat, slots.flatten()])
        _engine._all_reduce(buf, None)                        # the product's collective helper: ONE exchange
        flat, slots = buf[:flat.numel()], buf[flat.numel():].view(2, world, -1)
        stats_s, stats_t = _unpack(flat, sizes, len(layers), len(te), work.d_student, work.d_teacher)
        rows_local = local_batch * work.n_student
        rows = world * rows_local
        merged, o = [], 0
        for g, _ in stats_s + stats_t:
            d = g.shape[0]
            d_r, mu_r = slots[0, :, o:o + d].double(), slots[1, :, o:o + d].double()
            o += d
            mu = mu_r.mean(0)
            delta = mu_r - mu
            g = g.double() + sum(torch.outer(d_r[r], delta[r]) + torch.outer(delta[r], d_r[r])
                                 + rows_local * torch.outer(delta[r], delta[r]) for r in range(world))
            c = (d_r + rows_local * delta).sum(0)
            # back to the unshifted frame the CPU model takes
            merged.append(((g + torch.outer(mu, c) + torch.outer(c, mu) + rows * torch.outer(mu, mu)).float(),
                           (c + rows * mu).float()))
        stats_s, stats_t = merged[:len(layers)], merged[len(layers):]
        proj_s, proj_t, logt = cs.selector_state(work)
        sel = km.selector_model(stats_s, stats_t, rows, rows, proj_s, proj_t, logt)
        if rank == 0:
            out.put((sel["ranks"], torch.stack(sel["weights"]).tolist()))
    finally:
        dist.destroy_process_group()


def test_sharded_statistics_equal_concatenated_batch():
    world, local_batch = 2, 8
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, local_batch, out)) for r in range(world)]
    for p in procs:
        p.start()
    ranks, weights = out.get()
    weights = torch.tensor(weights)
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    work = cs.workload("c1", world * local_batch)
    _, _, st, te, _ = syn.make_inputs(work, seed=4)
    layers = sorted(st)
    proj_s, proj_t, logt = cs.selector_state(work)
    rows = work.batch * work.n_student
    ref = km.selector_model([km.token_stats(st[l]) for l in layers], [km.token_stats(te[k]) for k in sorted(te)],
                            rows, rows, proj_s, proj_t, logt)
    assert ranks == ref["ranks"]
    assert (weights - torch.stack(ref["weights"])).abs().max() < 1e-5


def test_shards_are_slices_of_the_global_batch():
    work8, work16 = cs.workload("c1", 8), cs.workload("c1", 16)
    full = syn.make_inputs(work16, seed=4)
    lo = syn.make_inputs(work8, seed=4, batch_offset=0)
    hi = syn.make_inputs(work8, seed=4, batch_offset=8)
    for k in full[2]:
        assert torch.equal(torch.cat([lo[2][k], hi[2][k]]), full[2][k])
    for k in full[3]:
        assert torch.equal(torch.cat([lo[3][k], hi[3][k]]), full[3][k])
        assert torch.equal(torch.cat([lo[4][k], hi[4][k]]), full[4][k])
    assert torch.equal(torch.cat([lo[1], hi[1]]), full[1])


def _shard_worker(rank, world, port, q, d, out):
    import torch.distributed as dist
    from basd_b200 import _engine as eng
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(7)
    kmats = torch.randn(q, d, d, generator=g)                       # identical on every rank
    calls = []

    def solver(k):                                                  # stands in for sym_eig (CUDA)
        calls.append(k.shape[0])
        return k.diagonal(dim1=1, dim2=2).contiguous() * 3.0, (k * 2.0 + 1.0).contiguous()

    lam, vt = eng.sharded_sym_eig(kmats, None, world, solver=solver)
    ok = bool(torch.equal(lam, kmats.diagonal(dim1=1, dim2=2) * 3.0) and torch.equal(vt, kmats * 2.0 + 1.0))
    if rank == 0:
        out.put((ok, calls))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,q", [(2, 16), (3, 16), (4, 5)])
def test_selector_eigenproblems_are_dealt_to_ranks_and_gathered_back(world, q):
    """Every rank solves ceil(q / world) of the q identical problems; the all-gather returns them
    in problem order (the ranks hold bit-identical statistics)."""
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_shard_worker, args=(r, world, port, q, 6, out)) for r in range(world)]
    for p in procs:
        p.start()
    ok, calls = out.get()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    assert ok
    assert calls == [(q + world - 1) // world]
