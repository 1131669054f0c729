"""The multi-GPU exchange of the C ABI (sfb_comm_unique_id / sfb_comm_init / sfb_allgather_lnL): one ncclAllGather of
the lnL shard per step, issued by the library — no torch collective.  A single-rank communicator runs on any GPU
box; the two-rank test needs two devices and hands the NCCL unique id over a multiprocessing queue (no
torch.distributed anywhere: the host side could be any language)."""
import numpy as np
import pytest

from starfish_b200 import synth

pytestmark = pytest.mark.gpu


def test_single_rank_communicator_roundtrip():
    import torch

    from starfish_b200.engine import LikelihoodEngine

    eng = LikelihoodEngine(256, 0, 1, 4)
    eng.comm_init(0, 1, eng.comm_unique_id())
    local = torch.arange(4, dtype=torch.float64, device="cuda") * 0.5 - 1.0
    out = eng.allgather_lnl(local)
    torch.cuda.synchronize()
    assert torch.equal(out, local)
    eng.close()


def _rank_main(rank, world, uid_q, out_q):
    import torch

    from starfish_b200.engine import LikelihoodEngine

    torch.cuda.set_device(rank)
    N, B = 384, 6
    per = B // world
    d = synth.stage_inputs_direct(N, per, n_comp=0, n_local=0, first_walker=rank * per)
    eng = LikelihoodEngine(N, 0, 1, per, device=rank)
    eng.set_data(d["wave"], d["sigma"], d["data_flux"])
    if rank == 0:
        uid = eng.comm_unique_id()
        for _ in range(world - 1):
            uid_q.put(uid)
    else:
        uid = uid_q.get(timeout=120)
    eng.comm_init(rank, world, uid)
    lnL, info = eng.log_likelihood(None, None, d["model_flux"], glob=d["glob"])
    full = eng.allgather_lnl(lnL)
    torch.cuda.synchronize()
    out_q.put((rank, full.cpu().numpy().copy(), lnL.cpu().numpy().copy()))
    eng.close()


def test_two_rank_allgather_of_lnl_shards():
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    ctx = mp.get_context("spawn")
    uid_q, out_q = ctx.Queue(), ctx.Queue()
    procs = [ctx.Process(target=_rank_main, args=(r, 2, uid_q, out_q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict()
    for _ in range(2):
        r, full, local = out_q.get(timeout=300)
        res[r] = (full, local)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = np.concatenate([res[0][1], res[1][1]])
    assert np.array_equal(res[0][0], expect) and np.array_equal(res[1][0], expect)
