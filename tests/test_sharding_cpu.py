"""CPU tests of the N>1 path: shard arithmetic, and a world_size-2 gloo run in which each rank reduces its own shard of
a rollout statistic with the oracle (standing in for the GPU) and the final gather reproduces the single-rank result."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from hierarchicalkarting_b200 import sharding

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_ranges_partition_everything():
    for total in (0, 1, 7, 65536, 1_000_003):
        for world in (1, 2, 3, 4, 8):
            spans = [sharding.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
            offs = [sharding.rollout_shard(total, r, world) for r in range(world)]
            assert sum(c for _, c in offs) == total
    with pytest.raises(ValueError):
        sharding.shard_range(10, 2, 2)


WORKER = r'''
import os, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, os.environ["HK_ROOT"])
from hierarchicalkarting_b200 import sharding, tracks
from oracle import oracle as O
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
track = tracks.COMPLEX
g = O.Game(track.sections_array(), track.n_sections, tracks.kart_array(2), 2, tracks.game_params(track))
leaf = tracks.root_state(track, 3, [2, 3], teams=[0, 1], tire_age=2500, times=[0, 80])
total = 3001
off, cnt = sharding.rollout_shard(total, rank, world)
part = g.rollouts(leaf, cnt, mode=0, seed=20260003, rollout_offset=off)
summary = np.concatenate([part["visit"].astype(np.float64), part["reward_sum"].ravel(), [part["plies"]]])
tot = sharding.gather_sum(dist, summary)
t = sharding.max_over_ranks(dist, 1.0 + rank)
# the sequential tree search shards over trees: a rank passes seed + first tree, so every tree keeps the key of the one-rank call
import ctypes
from oracle import structs as OS
n_trees, iters, seed = 23, 40, 777
roots = np.zeros(n_trees, dtype=OS.GAME_STATE_DTYPE)
for r in range(n_trees):
    st = tracks.root_state(track, r, [1 + r % 4, 1 + (r + 1) % 4], teams=[0, 1], tire_age=2500, times=[0, 5 * r])
    roots[r] = np.frombuffer(bytes(st), dtype=OS.GAME_STATE_DTYPE)[0]
lo, cnt_t, seed_r = sharding.tree_shard(n_trees, seed, rank, world)
mine = O.tree_search_batch(g, roots[lo:lo + cnt_t], iters, seed=seed_r, mode=0, threads=1)
digest = np.zeros(n_trees)
digest[lo:lo + cnt_t] = [float(np.frombuffer(mine["best"][k].tobytes(), dtype=np.uint8).astype(np.float64).sum() + 1000.0 * mine["n_nodes"][k]) for k in range(cnt_t)]
digest_all = sharding.gather_sum(dist, digest)
if rank == 0:
    one = O.tree_search_batch(g, roots, iters, seed=seed, mode=0, threads=1)
    want = [float(np.frombuffer(one["best"][k].tobytes(), dtype=np.uint8).astype(np.float64).sum() + 1000.0 * one["n_nodes"][k]) for k in range(n_trees)]
    assert np.array_equal(digest_all, np.array(want)), "sharded tree searches differ from the one-rank call"
if rank == 0:
    whole = g.rollouts(leaf, total, mode=0, seed=20260003, rollout_offset=0)
    ref = np.concatenate([whole["visit"].astype(np.float64), whole["reward_sum"].ravel(), [whole["plies"]]])
    assert np.array_equal(tot[:36], ref[:36]) and np.allclose(tot, ref, rtol=1e-12), "sharded statistics differ"
    assert t == float(world)
    print("SHARD_OK")
dist.destroy_process_group()
'''


def test_two_rank_gloo_gather(tmp_path):
    from oracle import oracle as O
    O.build()
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    env = dict(os.environ, HK_ROOT=ROOT, OMP_NUM_THREADS="1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", str(port), str(script)], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert out.returncode == 0 and "SHARD_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-3000:]
