"""GPU: the sharded supervised path through NCCL on the one GPU the test box has (a 1-rank group: same
code path -- asynchronous operand all-gather, local bank shard first, X all-gather -- without peers),
against the single-GPU path.  Runs scripts/check_supervised_sharded.py in a subprocess so that the NCCL
communicator does not live in the pytest process; the 2-rank run of the same script is recorded in
profiles/r01_check_supervised_sharded_n2.log.  (Named zz: runs after the parity tests proper.)"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.timeout(600)
def test_supervised_sharded_path_single_rank_nccl():
    env = dict(os.environ, RANK="0", WORLD_SIZE="1", LOCAL_RANK="0", MASTER_ADDR="127.0.0.1",
               MASTER_PORT=str(29600 + os.getpid() % 300))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "check_supervised_sharded.py")], env=env, cwd=ROOT,
                       capture_output=True, text=True, timeout=500)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("supervised:")]
    assert len(lines) == 4 and all(ln.endswith("OK") for ln in lines), r.stdout
