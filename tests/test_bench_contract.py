"""bench.py contract checks that need no GPU: the reference (CPU) arm prints one JSON line with the
required keys, and the synthetic prompt has the documented shape at every N."""
import json
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "prefill_tokens_per_sec" and d["unit"] == "tokens/s"
    assert d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["config"]["tokens"] == 18432


def test_prompt_shape_is_the_same_for_every_rank_count():
    from long_vita_b200.config import LongVITAConfig
    from long_vita_b200.synthetic import build_prompt

    cfg = LongVITAConfig.long_vita_14b()
    ids, idx = build_prompt(cfg, 64, 16, pad_multiple=2048)
    assert ids.shape == (1, 18432) and idx.shape == (2, 64, 256)
    for cp in (1, 2, 4, 8):
        assert ids.shape[1] % (2 * cp * 128) == 0
    # frame layout: [VID_START] + 256 x [VID_CONTEXT] + [VID_END] (tools/inference_long_vita.py:730-748)
    assert torch.equal(idx[1, 0], torch.arange(1, 257)) and torch.equal(idx[1, 1], torch.arange(259, 515))
    assert (idx[0] == 0).all()
    ids128, idx128 = build_prompt(cfg, 512, 16, pad_multiple=2048)
    assert ids128.shape[1] == 133120
