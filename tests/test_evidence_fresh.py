"""The ncu counters behind bench.py's roofline block (profiles/traffic.json: DRAM bytes and warp instructions of the frame kernel on the
bench command) belong to ONE build of the kernels: the file carries the hash of the CUDA sources it was captured from and bench.py
refuses it on a mismatch (roofline.traffic / issue_frac / winstr_per_frame would then be null).  This test keeps the committed capture
and the committed sources together: change a kernel -> re-run scripts/measure_traffic.sh on the GPU and scripts/summarize_profiles.py."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_committed_ncu_counters_match_the_committed_kernel_sources():
    import bench
    t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    assert t["kernel_source_hash"] == bench.kernel_source_hash(), "profiles/traffic.json was captured from other kernel sources"
    assert t["frame_kernel_warp_instructions_per_launch"] > 1e9 and t["frame_kernel_dram_bytes_per_launch"] >= t["read_bytes"] > 1.8e8


def test_bench_lines_of_the_round_are_committed_and_consistent():
    """The bench lines profiles/README.md is generated from: the contract's keys are there, the fleet block scales, every frame of every
    workload ended `updated`."""
    lines = {n: json.load(open(os.path.join(ROOT, "profiles", "bench_r02_final_n%d.json" % n))) for n in (1, 2, 4, 8)}
    for n, d in lines.items():
        for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
                  "data", "config", "roofline", "e2e", "gpu_launches", "clocks", "fleet"):
            assert k in d, (n, k)
        assert d["n_gpus"] == n and d["vs_baseline"] is None and d["roofline"]["bound"] == "hbm"
        h = d["config"]["status_hist"]
        assert h["updated"] == h["frames"] and h["overflow"] == 0
        assert d["fleet"]["status_hist"]["updated"] == 23201
        assert not d["clocks"]["reasons"]
    assert lines[8]["fleet"]["value"] / lines[1]["fleet"]["value"] >= 7.0          # north star: >= 7x strong scaling at 8 GPUs
    assert lines[1]["value"] >= 1.0e5                                               # north star: >= 1e5 frames/s/GPU
    assert lines[1]["cpu_baseline"]["kind"] == "reference"
