"""CPU suite: the N > 1 host logic of bench.py (weak scaling by batch shard, max-over-ranks timing, rank-0 JSON)
exercised with world_size = 2 over gloo.  No GPU and no data-path collective is involved: the codec path shards by
image, the only exchange is the timing reduction."""
import json
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import json, os, sys
    sys.path.insert(0, %r)
    import torch, torch.distributed as dist
    import bench
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo")
    # each rank "measures" a different step time; the job time must be the slowest one
    local_ms = [100.0 + 50.0 * rank, 300.0 - 20.0 * rank]
    ms, e2e = bench.max_over_ranks(local_ms, torch.device("cpu"), world)
    val = bench.shard_throughput(10.0, ms, world)
    # shards are disjoint: rank r owns images [r * B, (r + 1) * B)
    B = 4
    mine = list(range(rank * B, (rank + 1) * B))
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    if rank == 0:
        print(json.dumps({"ms": ms, "e2e": e2e, "value": val, "shards": gathered}))
    dist.destroy_process_group()
""")


def test_world_size_two_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29617")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29617", str(script)],
                         capture_output=True, text=True, env=env, timeout=240)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    assert d["ms"] == 150.0 and d["e2e"] == 300.0
    assert abs(d["value"] - 2 * 10.0 / 0.150) < 1e-9
    assert d["shards"] == [[0, 1, 2, 3], [4, 5, 6, 7]]


def test_reference_arm_contract_fields():
    """--impl reference prints the same metric / unit and the cpu_baseline + e2e objects (checked on a tiny sample
    here only when the reference install exists; the arm itself runs on the GPU box's host cores)."""
    from oracle import oracle as orc

    if not orc.have_ref():
        import pytest

        pytest.skip("oracle/_ref not built")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--ref-sample", "1"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert d["impl"] == "reference" and d["unit"] == "MP/s" and d["cpu_baseline"]["kind"] == "reference"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["value"] > 0
