"""N > 1 host logic on CPU: two `gloo` ranks build host-only plans of the same Domain, take their row blocks
(fem2d_plan_row_blocks), and the rank-local slices reassemble into the full result.  The numeric values are stand-ins taken
from the oracle (no GPU here); what is under test is the sharding + the collective plumbing bench.py uses."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import recipes


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, recipe, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import fem_2d_b200 as F
        import oracle as O
        mo, mf = recipes.build_pair(recipe)
        df = F.Domain.from_mesh(mf)
        plan = F.Plan(df.view(), device=-1)
        bounds = plan.row_blocks(world)
        # all ranks must derive the same partition
        gathered = [torch.zeros(world + 1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(gathered, torch.from_numpy(bounds.astype(np.int64)))
        for g in gathered:
            assert torch.equal(g, gathered[0])
        s0, s1 = int(bounds[rank]), int(bounds[rank + 1])
        rows, cols = plan.pattern()
        if 0 < s0 < plan.nnz:
            assert rows[s0] != rows[s0 - 1]                      # blocks start at row boundaries
        # stand-in numeric phase: this rank's slice of the oracle result
        ref = O.galerkin_sample_gep_hcurl(O.Domain.from_mesh(mo), [4, 4])
        assert np.array_equal(ref.rows, rows) and np.array_equal(ref.cols, cols)
        mine_a = torch.from_numpy(ref.a[s0:s1].copy())
        # gather the slices on every rank (sizes differ: pad to the longest block)
        longest = int(np.max(np.diff(bounds.astype(np.int64))))
        buf = torch.zeros(longest, dtype=torch.float64)
        buf[: s1 - s0] = mine_a
        parts = [torch.zeros(longest, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(parts, buf)
        full = torch.cat([parts[r][: int(bounds[r + 1] - bounds[r])] for r in range(world)])
        assert np.array_equal(full.numpy().view(np.uint64), ref.a.view(np.uint64))
        # two-level partition (Elem-type rows + edge-type rows per rank): consistent across ranks, covers every slot exactly once
        b1, b2 = plan.row_blocks_split(world)
        g2 = [torch.zeros(2 * (world + 1), dtype=torch.int64) for _ in range(world)]
        dist.all_gather(g2, torch.from_numpy(np.concatenate([b1, b2]).astype(np.int64)))
        for g in g2:
            assert torch.equal(g, g2[0])
        assert b1[0] == 0 and b1[-1] == b2[0] and b2[-1] == plan.nnz
        cover = np.zeros(plan.nnz, dtype=np.int32)
        for r in range(world):
            cover[int(b1[r]):int(b1[r + 1])] += 1
            cover[int(b2[r]):int(b2[r + 1])] += 1
        assert np.all(cover == 1)
        if 0 < b2[0] < plan.nnz:
            assert rows[int(b2[0])] != rows[int(b2[0]) - 1]
        # max-over-ranks timing reduction as in bench.py
        t = torch.tensor([float(rank + 1)], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        assert t.item() == float(world)
        dist.barrier()
        open(os.path.join(out_dir, f"ok{rank}"), "w").write(f"{s0} {s1}")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("recipe", ["slepc", "cfg4_small"])
def test_two_rank_row_block_sharding(recipe, tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), recipe, str(tmp_path)), nprocs=world, join=True)
    spans = [tuple(map(int, open(tmp_path / f"ok{r}").read().split())) for r in range(world)]
    assert spans[0][0] == 0 and spans[0][1] == spans[1][0] and spans[1][1] > spans[1][0]
