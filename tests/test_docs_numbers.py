"""The headline numbers quoted in README.md are the ones in the committed bench lines under profiles/."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(name):
    with open(os.path.join(ROOT, "profiles", name)) as f:
        return json.loads([l for l in f.read().splitlines() if l.startswith("{")][-1])


def _k(x):
    return f"{x / 1000:.1f} k"


def test_readme_quotes_the_committed_bench_lines():
    readme = open(os.path.join(ROOT, "README.md")).read()
    d = _line("bench_r02_default.json")
    assert d["config"]["n_total"] == 4096 and d["config"]["t"] == 2731 and d["n_gpus"] == 1
    assert f"MODP {_k(d['value'])} verified shares/s" in readme
    assert f"{d['roofline']['frac']:.3f} of the measured 32-bit IMAD peak" in readme
    assert f"secp256k1 {d['also']['secp256k1']['value'] / 1000:.0f} k/s" in readme
    assert f"ristretto255 {_line('bench_r02_ristretto255.json')['value'] / 1000:.0f} k/s" in readme
    weak = [_line(f"bench_r02_n{n}.json") for n in (2, 4, 8)]
    assert " / ".join(_k(w["value"]) for w in weak) + " MODP shares/s" in readme
    for w, n in zip(weak, (2, 4, 8)):
        assert w["n_gpus"] == n and w["scaling"] == "weak" and w["config"]["n_total"] == n * 4096
    strong = [_line("bench_r02_n2.json"), _line("bench_r02_n4.json"), _line("bench_r02_n8_strong.json")]
    assert " / ".join(_k(s["also"]["strong"]["value"]) for s in strong) in readme
    c5 = weak[2]["also"]["c5"]
    assert c5["workload"].startswith("modp n=65536 t=43691") and f"{c5['ms_per_step'] / 1000:.1f} s per verification pass" in readme


def test_bench_lines_carry_the_contract_keys():
    for name in ("bench_r02_default.json", "bench_r02_ristretto255.json", "bench_r02_n2.json", "bench_r02_n8.json"):
        d = _line(name)
        for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                    "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline"):
            assert key in d, (name, key)
        assert d["steps"] >= 3 and d["warmup"] >= 3 and d["gpu_launches"] > 0
        assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
        assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(d["e2e"])
        assert {"bound", "achieved", "peak", "unit", "frac"} <= set(d["roofline"])
        assert abs(d["roofline"]["frac"] - d["roofline"]["achieved"] / d["roofline"]["peak"]) < 1e-9
    ref = _line("bench_r02_reference_arm.json")
    assert ref["impl"] == "reference" and ref["cpu_baseline"]["kind"] == "port" and ref["e2e"]["h2d_bytes_per_step"] == 0
    d = _line("bench_r02_default.json")
    assert {"value", "unit", "cores", "kind", "sample"} <= set(d["cpu_baseline"])
