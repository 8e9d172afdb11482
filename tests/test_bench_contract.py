"""CPU: the committed bench line (profiles/r2_final_bench_1gpu.json, written by `python bench.py --gpus 1 --steps 20 --warmup 5`
on a B200) carries every key of the measurement contract, and bench.py still produces those keys (static check of the writer)."""
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_committed_bench_line_has_the_contract_keys():
    d = json.load(open(os.path.join(ROOT, "profiles", "r2_final_bench_1gpu.json")))
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "clocks", "e2e", "roofline", "cpu_baseline", "gpu_launches", "parity", "extra_configs"):
        assert key in d, key
    assert d["dtype"] == "f64" and d["data"] == "synthetic" and d["vs_baseline"] is None and "workload" in d["config"]
    assert set(("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step")) <= set(d["e2e"])
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["value"] < d["value"]          # host buffers inside the timed region
    r = d["roofline"]
    assert set(("bound", "achieved", "peak", "unit", "frac", "traffic")) <= set(r)
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and "l2" in r                 # + the L2 gather roof
    assert set(("value", "unit", "cores", "kind", "sample")) <= set(d["cpu_baseline"]) and d["cpu_baseline"]["kind"] in ("port", "reference")
    assert set(("sm_mhz", "sm_max_mhz", "reasons")) <= set(d["clocks"]) and not d["clocks"]["reasons"]
    assert d["gpu_launches"] > 0
    assert d["parity"]["ok"] and d["parity"]["max_rel_err"] <= d["parity"]["tolerance"] == 1e-4
    for c in ("C3", "C4", "C5"):                                                            # every BASELINE configuration, full size
        e = d["extra_configs"][c]
        assert e["parity"]["ok"] and e["ms_per_step"] > 0 and "roofline" in e


def test_bench_writer_emits_the_contract_keys():
    src = open(os.path.join(ROOT, "bench.py")).read()
    for key in ('"metric"', '"value"', '"unit"', '"n_gpus"', '"ms_per_step"', '"higher_is_better"', '"scaling"', '"vs_baseline"',
                '"dtype"', '"data"', '"config"', '"clocks"', '"e2e"', '"roofline"', '"gpu_launches"', '"cpu_baseline"',
                '"h2d_bytes_per_step"', '"d2h_bytes_per_step"', '"impl": "reference"'):
        assert key in src, key
    assert re.search(r"add_argument\(\"--gpus\".*default=1", src) and "--steps" in src and "--warmup" in src and "--impl" in src
