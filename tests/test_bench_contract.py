"""The driver-facing contract of bench.py: one JSON line with the agreed keys.  The reference arm (CPU restatement) runs
here without a GPU at a reduced size; the GPU arm is checked on the B200 box at a reduced size (-m gpu)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline"}


def _run(args):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout
    return json.loads(lines[0])


def test_reference_arm_line():
    d = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--log-n", "12"])
    assert d["impl"] == "reference" and BASE_KEYS <= set(d)
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert d["config"]["workload"].startswith("stark_proof_2^12") and d["vs_baseline"] is None and d["higher_is_better"] is True


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--log-n", "12"],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0 and out.stdout.strip() == ""


@pytest.mark.gpu
def test_gpu_arm_line():
    d = _run(["--steps", "2", "--warmup", "3", "--log-n", "14", "--orders", "4096"])
    assert BASE_KEYS | {"roofline", "clocks", "gpu_launches", "proof_gen_s", "stage_ms", "proof_sha256", "aux"} <= set(d)
    assert len(d["proof_sha256"]) == 64
    aux = d["aux"]
    assert "error" not in aux, aux
    for k in ("cfg0_pedersen_1024", "cfg1_ntt_2^18_single", "cfg1_ntt_2^18_batch64", "cfg4_orders_valid_mix", "cfg4_orders_invalid_mix"):
        assert aux[k]["ms"] > 0 and 0 < aux[k]["frac_of_int_ceiling"] < 1, k
    assert aux["cfg0_pedersen_1024"]["oracle_sample_ok"] and aux["cfg4_orders_valid_mix"]["statuses_as_expected"]
    if "cpu_reference" in aux["cfg0_pedersen_1024"]:
        assert aux["cfg0_pedersen_1024"]["cpu_reference"]["kind"] == "reference"
        assert aux["cfg0_pedersen_1024"]["cpu_reference"]["gpu_equals_reference"]
        assert aux["cfg4_orders_valid_mix"]["cpu_reference"]["gpu_equals_reference"]
    assert d["gpu_launches"] > 0 and d["verified_by_oracle"] is True and d["scaling"] == "strong"
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert d["e2e"]["h2d_bytes_per_step"] == 25 * (1 << 14) * 32 and d["e2e"]["d2h_bytes_per_step"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and "sample" in d["cpu_baseline"]
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
