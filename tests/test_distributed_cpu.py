"""World-size-2 `gloo` test (CPU) of the host-side multi-GPU logic: each rank owns a row shard
(+ halo), produces a first-fail vector, and ONE all-reduce(MIN) (+SUM of counts) gives the same
result as checking the whole circuit.  The per-shard checks run through the CPU oracle here (no
GPU in this container); the -m gpu tests run the same sharding through the CUDA kernels."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent(
    """
    import os, sys
    sys.path.insert(0, os.path.join(%r, "tests")); sys.path.insert(0, %r)
    import numpy as np, torch, torch.distributed as dist
    import oracle_lib
    from zkevm_specs_b200 import synth
    from zkevm_specs_b200.evm_circuit.table import fixed_table_matrix
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    w = synth.evm_trace(32, seed=11)
    S = w["steps"].copy()
    S[9, 70, 0] += np.uint64(1)          # gas corruption in rank 1's shard
    S[1, 5, 0] += np.uint64(1)           # rw_counter corruption in rank 0's shard
    fixed = fixed_table_matrix()
    n = S.shape[1] - 1
    lo, hi = rank * n // world, (rank + 1) * n // world
    shard = np.ascontiguousarray(S[:, lo : hi + 1])      # +1 halo step
    ff, fc = oracle_lib.check_evm(shard, w["bytecode"], w["rw"], fixed, 0, hi - lo, lo, 0)
    # uint32 first_fail with 0xFFFFFFFF = pass: MIN over uint32 == MIN over (x ^ 0x80000000) as int32
    t = torch.from_numpy(ff.view(np.int32).copy()) ^ -0x80000000
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    ff_all = (t ^ -0x80000000).numpy().view(np.uint32)
    c = torch.from_numpy(fc.astype(np.int64)); dist.all_reduce(c, op=dist.ReduceOp.SUM)
    if rank == 0:
        whole, wc = oracle_lib.check_evm(S, w["bytecode"], w["rw"], fixed)
        assert np.array_equal(ff_all, whole), (ff_all[ff_all != whole], whole[ff_all != whole])
        assert np.array_equal(c.numpy().astype(np.uint64), wc)
        assert (whole != 0xFFFFFFFF).sum() >= 2
        print("OK")
    dist.destroy_process_group()
    """
)


def test_row_sharded_allreduce_min_equals_whole(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % (ROOT, ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                         capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0 and "OK" in out.stdout, out.stdout[-2000:] + out.stderr[-3000:]
