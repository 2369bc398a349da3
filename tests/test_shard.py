"""CPU: the multi-GPU sharding logic (byte-balanced split by sequence index, no collective).

The N > 1 path is covered with a real 2-process ``gloo`` group on CPU: each rank takes its shard
with bioseq_b200.shard, runs the *oracle* on it (there is no GPU here; the GPU flavour of the same
check is tests/test_gpu_parity.py::test_sharded_equals_single_device), the shards are gathered
and must equal the unsharded oracle output; timing reduction (max over ranks) is exercised too."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from bioseq_b200.shard import shard_bounds, take_shard
from helpers import gen

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_bounds_balance_and_cover():
    buf, offs = gen(5, 10_000, 0, 900, b"ACGT")
    for world in (1, 2, 3, 4, 8):
        b = shard_bounds(offs, world)
        assert b[0] == 0 and b[-1] == 10_000 and len(b) == world + 1 and np.all(np.diff(b) >= 0)
        sizes = np.array([offs[b[r + 1]] - offs[b[r]] for r in range(world)])
        assert sizes.sum() == offs[-1]
        assert sizes.max() - sizes.min() <= 2 * 900          # balanced to within ~one sequence
        cat = np.concatenate([take_shard(buf, offs, b, r)[0] for r in range(world)])
        assert np.array_equal(cat, buf)
        for r in range(world):
            sb, so = take_shard(buf, offs, b, r)
            assert so[0] == 0 and so[-1] == sb.size and len(so) == b[r + 1] - b[r] + 1
    b = shard_bounds(offs, 4, align=128)
    assert all(x % 128 == 0 for x in b[1:-1])


def test_bounds_edge_cases():
    assert shard_bounds(np.array([0]), 4).tolist() == [0, 0, 0, 0, 0]                 # empty batch
    assert shard_bounds(np.array([0, 0, 0, 0]), 2).tolist()[0::2] == [0, 3]           # only empty sequences
    b = shard_bounds(np.array([0, 1000, 1001, 1002]), 2)                               # one dominant sequence
    assert b[0] == 0 and b[-1] == 3
    b = shard_bounds(np.array([0, 5, 9]), 8)                                          # more ranks than sequences
    assert b[0] == 0 and b[-1] == 2 and np.all(np.diff(b) >= 0)
    with pytest.raises(ValueError):
        shard_bounds(np.array([0, 1]), 0)


WORKER = r'''
import os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
from bioseq_b200.shard import shard_bounds, take_shard
from oracle.oracle import OracleTokenizer
from helpers import gen
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
buf, offs = gen(77, 3001, 0, 300, b"ACDEFGHIKLMNPQRSTVWYX")
tok = OracleTokenizer("PROTEIN", bos=True, eos=True, padchar=True)
bounds = shard_bounds(offs, world)
sb, so = take_shard(buf, offs, bounds, rank)
t0 = time.perf_counter()
mine = tok.batch_tokenize((np.ascontiguousarray(sb), np.ascontiguousarray(so)), padlen=304, batch_first=True)
dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64)
dist.barrier()
dist.all_reduce(dt, op=dist.ReduceOp.MAX)                      # bench.py reduces timings the same way
parts = [None] * world
dist.all_gather_object(parts, mine)
if rank == 0:
    full = tok.batch_tokenize((buf, offs), padlen=304, batch_first=True)
    assert np.array_equal(np.concatenate(parts, axis=0), full)
    sf = [tok.batch_tokenize(take_shard(buf, offs, bounds, r), padlen=304, batch_first=False) for r in range(world)]
    assert np.array_equal(np.concatenate(sf, axis=1), full.T)
    assert float(dt) > 0
    print("SHARD_OK", world, [int(p.shape[0]) for p in parts])
dist.destroy_process_group()
'''


def test_two_rank_gloo_shards_concatenate_to_full_result(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script), ROOT], env=dict(env, RANK=str(r), LOCAL_RANK=str(r)),
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "SHARD_OK 2" in outs[0]
