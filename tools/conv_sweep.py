#!/usr/bin/env python
"""Time individual conv layers of the FFHQ UNet step under the tuning knobs of conv_tc.cu.

The knobs are read once per process (static getenv), so this script re-executes itself per configuration:
    python tools/conv_sweep.py                 # all configurations below, markdown table on stdout
    AFLDM_TC_HALO=0 python tools/conv_sweep.py --one   # one process, one line of timings (us per launch)
"""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

LAYERS = [  # B, H, Cin, Cout, k
    (16, 32, 192, 192, 3), (16, 32, 384, 192, 3), (16, 32, 576, 192, 3), (16, 32, 384, 384, 3),
    (16, 16, 384, 384, 3), (16, 16, 768, 384, 3), (16, 16, 192, 384, 3),
    (16, 32, 192, 576, 1), (16, 16, 384, 1152, 1), (16, 8, 384, 384, 3), (16, 4, 768, 768, 3),
]
CONFIGS = [
    {"AFLDM_TC_HALO": "0"},
    {},
    {"AFLDM_TC_TWO": "0"},
    {"AFLDM_TC_TWO": "0", "AFLDM_TC_BN": "64"},
    {"AFLDM_TC_TWO": "0", "AFLDM_TC_BN": "96"},
    {"AFLDM_TC_TWO": "0", "AFLDM_TC_BN": "128"},
    {"AFLDM_TC_TWO": "0", "AFLDM_TC_BN": "192"},
    {"AFLDM_TC2_BN": "64"},
    {"AFLDM_TC2_BN": "96"},
    {"AFLDM_TC2_BN": "128"},
    {"AFLDM_TC2_BN": "192"},
    {"AFLDM_TC_STAGES": "2"},
    {"AFLDM_TC_STAGES": "3"},
]


def one():
    import torch
    from afldm_b200 import ops
    g = torch.Generator().manual_seed(0)
    out = []
    for (b, h, cin, cout, k) in LAYERS:
        x = torch.randn(b, h, h, cin, generator=g).cuda()
        w = ops.pack_conv_weight((torch.randn(cout, cin, k, k, generator=g) * 0.02).cuda())
        bias = torch.randn(cout, generator=g).cuda()
        y = torch.empty(b, h, h, cout, device="cuda")
        for _ in range(3):
            ops.conv2d(x, w, bias, k, algo="tf32", out=y)
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            for _ in range(20):
                ops.conv2d(x, w, bias, k, algo="tf32", out=y)
        gr.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e9
        for _ in range(3):
            e0.record()
            gr.replay()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) * 1000 / 20)
        out.append(best)
    print("RESULT " + " ".join(f"{t:.1f}" for t in out))


def main():
    if "--one" in sys.argv:
        return one()
    print("| config | " + " | ".join(f"{h}x{h} {ci}->{co} k{k}" for (_, h, ci, co, k) in LAYERS) + " |")
    print("|---|" + "---|" * len(LAYERS))
    for cfg in CONFIGS:
        env = dict(os.environ)
        env.update(cfg)
        r = subprocess.run([sys.executable, __file__, "--one"], env=env, capture_output=True, text=True, timeout=300)
        line = [l for l in r.stdout.splitlines() if l.startswith("RESULT")]
        name = " ".join(f"{k[9:]}={v}" for k, v in cfg.items()) or "default"
        print(f"| {name} | " + (" | ".join(line[0].split()[1:]) if line else "FAILED " + r.stderr[-200:].replace("\n", " ")) + " |", flush=True)


if __name__ == "__main__":
    main()
