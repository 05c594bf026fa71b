"""k_generations has two CTA shapes — 512 threads (shards below ZZ_GEN_LARGE_MIN genes, batched chains) and 640 threads (above).  The
parity tests use small gene sets and therefore the first; this one repeats the chain-replay tests (oracle chain decision for
decision, golden trajectory, any-grid-size sweeps) in a subprocess with ZZ_GEN_LARGE_MIN=1, i.e. through the 640-thread shape."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.timeout(900)
def test_chain_replay_through_the_640_thread_shape():
    env = dict(os.environ, ZZ_GEN_LARGE_MIN="1")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_chain.py"), "-m", "gpu", "-q", "-x",
                        "-k", "match_the_oracle_chain or golden or run_sweeps_matches or tuning_point or sample_rows", "-p", "no:cacheprovider"],
                       capture_output=True, text=True, env=env, cwd=ROOT, timeout=800)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert " passed" in r.stdout and "failed" not in r.stdout
