"""N > 1 on CPU (gloo, world_size 2, 127.0.0.1): the tensor-parallel sharding arithmetic and where the exchange
steps sit, and the replica aggregation rule bench.py uses.

One process per rank, as in production; each rank runs ITS slice of the row-parallel projections
(oracle.llama_ref.LlamaOracle(tp=2, rank=r, allreduce=...)) and the two all-reduces per layer go through
torch.distributed.  The result must be bit-identical to the in-process simulation of both ranks that the GPU TP
tests compare the CUDA path against, and close to the unsharded model.
"""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import llama_ref as ref
    from oracle.weights import ModelDesc, SynthWeights
    calls = []

    def allreduce(partial16):
        t = torch.from_numpy(partial16.astype(np.float32))
        dist.all_reduce(t, op=dist.ReduceOp.SUM)       # fp16 partials, summed exactly in fp32, rounded once
        calls.append(partial16.shape)
        return t.numpy().astype(np.float16)

    for quant in (1, 0):
        desc = ModelDesc(256, 512, 2, 4, 2, 320, cache_layout=3, cache_mode=1, page_size=16, quant_method=quant, max_position=64)
        w = SynthWeights(desc, 0xB200)
        orc = ref.LlamaOracle(desc, w, 64, tp=world, rank=rank, allreduce=allreduce)
        rng = np.random.default_rng(3)
        prompts = [list(map(int, rng.integers(0, desc.vocab_size, n))) for n in (9, 4)]
        pages = [[0], [16]]
        l0 = orc.forward(ref.build_step(desc, prompts, [0, 0], 0, page_tables=pages))
        nxt = l0.argmax(axis=1)
        l1 = orc.forward(ref.build_step(desc, [[int(t)] for t in nxt], [9, 4], 2, page_tables=pages))
        np.savez(Path(out_dir) / f"rank{rank}_q{quant}.npz", l0=l0, l1=l1, calls=np.asarray(len(calls)))
    # replica aggregation of bench.py: value = units of all ranks / max over ranks of the device time
    ms = torch.tensor([10.0 + 5.0 * rank], dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    np.save(Path(out_dir) / f"max{rank}.npy", ms.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_tensor_parallel_two_ranks_gloo(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    from oracle import llama_ref as ref
    from oracle.weights import ModelDesc, SynthWeights
    for quant in (1, 0):
        r0 = np.load(tmp_path / f"rank0_q{quant}.npz")
        r1 = np.load(tmp_path / f"rank1_q{quant}.npz")
        # every rank ends with the same logits (inputs replicated, partials all-reduced: SURVEY 8e)
        assert np.array_equal(r0["l0"], r1["l0"]) and np.array_equal(r0["l1"], r1["l1"])
        # two exchange steps per layer per forward: 2 layers x 2 x 2 forwards (per quant mode, cumulative)
        assert int(r0["calls"]) == (8 if quant == 1 else 16)
        desc = ModelDesc(256, 512, 2, 4, 2, 320, cache_layout=3, cache_mode=1, page_size=16, quant_method=quant, max_position=64)
        w = SynthWeights(desc, 0xB200)
        rng = np.random.default_rng(3)
        prompts = [list(map(int, rng.integers(0, desc.vocab_size, n))) for n in (9, 4)]
        pages = [[0], [16]]
        sim = ref.LlamaOracle(desc, w, 64, tp=world)                     # both ranks simulated in one process
        s0 = sim.forward(ref.build_step(desc, prompts, [0, 0], 0, page_tables=pages))
        assert np.array_equal(s0, r0["l0"])
        s1 = sim.forward(ref.build_step(desc, [[int(t)] for t in s0.argmax(axis=1)], [9, 4], 2, page_tables=pages))
        assert np.array_equal(s1, r0["l1"])
        one = ref.LlamaOracle(desc, w, 64)                                 # unsharded model: same function
        u0 = one.forward(ref.build_step(desc, prompts, [0, 0], 0, page_tables=pages))
        assert np.abs(u0 - s0).max() <= (0.05 if quant else 5e-3) * np.abs(u0).max()
    assert float(np.load(tmp_path / "max0.npy")[0]) == 15.0 == float(np.load(tmp_path / "max1.npy")[0])
