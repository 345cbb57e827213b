"""Peer-memory exchange (xs_exchange_*): the sending end fused into the search's emit step (xs_search_dev_push),
the stand-alone push kernel, the per-query arrival flags the merge kernel waits on, the acknowledgements that free
a mailbox slot, and the certificate words that travel with the lists.  World 1 in-process; world 2 as two processes
that SHARE device 0 (gloo carries the handles, CUDA IPC maps the mailboxes) so that a one-GPU box still exercises
shard + exchange + merge; world 2 over NCCL under torchrun when two GPUs exist."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_exchange_single_rank(pkg, synth, oracle):
    import importlib
    import torch
    sharded = importlib.import_module(pkg.__name__ + ".sharded")
    v, q = synth.gaussian(6000, 40, d=128)
    index = pkg.ExactIndex(v.T, id_offset=1000)
    ex = sharded.PeerExchange(0, 40, 50)
    shard, searcher = sharded.make_searcher(index, 0, exchange=ex)
    qd = torch.from_numpy(np.ascontiguousarray(q.T)).cuda()
    ref_i, ref_s = oracle.topk_ip(v, q, 50)
    s64 = oracle.scores_f64(v, q)
    for epoch in range(5):                                         # both slots, several epochs, varying sizes
        nq, k = ((40, 50), (1, 50), (17, 3))[epoch % 3]
        ids, sims = searcher.search(qd[:nq].contiguous(), k)
        ids, sims = ids.cpu().numpy(), sims.cpu().numpy()
        for j in range(nq):
            ok, msg = oracle.compare_topk(ids[j] - 1000, ref_i[j, :k], lambda i, j=j: s64[i, j])
            assert ok, f"epoch {epoch} query {j}: {msg}"
        np.testing.assert_allclose(sims, ref_s[:nq, :k], rtol=1e-5, atol=1e-7)
    a, b = searcher.search_async(qd, 50), searcher.search_async(qd, 50)
    np.testing.assert_array_equal(a.result()[0].cpu().numpy(), b.result()[0].cpu().numpy())
    # the native pipeline without an exchange (single shard) and with one
    for ex_p in (None, sharded.PeerExchange(0, 40, 50)):
        _, pipe = sharded.make_searcher(index, 0, lanes=2, exchange=ex_p, pipeline=(40, 50))
        hs = [pipe.search_async(qd, 50), pipe.search_async(qd[:7].contiguous(), 9)]
        np.testing.assert_array_equal(hs[0].result()[0].cpu().numpy(), a.result()[0].cpu().numpy())
        np.testing.assert_array_equal(hs[1].result()[0].cpu().numpy(), a.result()[0].cpu().numpy()[:7, :9])
        with pytest.raises(ValueError):
            pipe.search(torch.zeros((41, 128), device="cuda"), 50)      # larger than the pipeline was sized for
        pipe.close()
        if ex_p is not None:
            ex_p.close()
    # protocol errors are refused at the ABI, not left to hang a kernel
    with pytest.raises(ValueError):
        ex.push(torch.empty(sharded.packed_bytes(41, 50), dtype=torch.uint8, device="cuda"), 41, 50, 0)   # more queries than the mailbox holds
    with pytest.raises(ValueError):
        ex.merge(40, 50, 1)                                         # a merge without its push
    packed = shard.local_search(qd, 50, 0)
    ex.push(packed, 40, 50, 0)                                      # the stand-alone push kernel
    with pytest.raises(ValueError):
        ex.push(packed, 40, 50, 0)                                  # the slot's previous result has not been merged
    with pytest.raises(ValueError):
        shard.local_push(qd, 50, ex, 0)
    got = ex.merge(40, 50, 0)
    np.testing.assert_array_equal(got[0].cpu().numpy(), a.result()[0].cpu().numpy())
    assert int(got[2].sum().item()) == 0
    ex.close()
    # a payload that takes several push CTAs per peer (2.3 MB) and the two-kernel finalise: 3000 queries, k = 64
    vq, _ = synth.gaussian(3000, 1, d=128)
    big = sharded.PeerExchange(0, 3000, 64)
    shard2, searcher = sharded.make_searcher(index, 0, exchange=big)
    qd = torch.from_numpy(np.ascontiguousarray(vq.T)).cuda()
    want_i, want_s = index.search(vq.T, 64)
    for _ in range(3):
        ids, sims = searcher.search(qd, 64)
        np.testing.assert_array_equal(ids.cpu().numpy(), want_i)
        np.testing.assert_array_equal(sims.cpu().numpy(), want_s)
    packed = shard2.local_search(qd, 64, 0)
    big.push(packed, 3000, 64, 1)
    np.testing.assert_array_equal(big.merge(3000, 64, 1)[0].cpu().numpy(), want_i)
    big.close()
    index.close()


def _torchrun(script, nproc, extra_env=None, timeout=900):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc), "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", script)]
    env = dict(os.environ)
    env.update(extra_env or {})
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT, env=env)


def test_two_shards_on_one_device():
    """Two ranks, ONE GPU: row-sharded search + peer exchange + merge against the oracle, on the families that need
    the exact re-run (crowded scores, duplicated rows) as well as the plain one -- runs on the driver's 1-GPU box."""
    r = _torchrun("exchange_check.py", 2, {"XS_CHECK_ONE_DEVICE": "1"})
    assert r.returncode == 0 and "exchange_check ok" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


def test_exchange_two_ranks():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run with gpurun --gpus 2)")
    r = _torchrun("exchange_check.py", 2)
    assert r.returncode == 0 and "exchange_check ok" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
