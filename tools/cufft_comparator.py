"""cuFFT (through torch.fft) as a speed comparator for the RIR-partition transform (BASELINE.json north_star: "cuFFT
serves only as a correctness and speed comparator"). Same batch as one benchmark step: all RIR partitions of 128
C3-style scenes, P = 2048 samples zero-padded to 4096, real-to-complex.

    python tools/cufft_comparator.py [scenes]

cuFFT needs the zero-padded blocks materialised (or a load callback); both variants are timed: transform only (input
already padded in HBM) and pad + transform. k_ir_fft reads the unpadded taps, also produces the per-partition energies
of normalize_irs and writes the spectra in the layout the multiply-accumulate kernel wants.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from audiblelight_b200 import workload as wl  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 128
P = 2048
dev = torch.device("cuda", 0)
n_part = 0
for si in range(S):
    sp = wl.c3_scene_spec(si, augment=True)
    K = -(-sp.n_ir_samples // P)
    n_part += sum(sp.channels * e.n_irs * K for e in sp.events)
print(f"{S} scenes: {n_part} partition transforms of {2 * P} real points ({n_part * P * 4 / 1e9:.2f} GB of taps)")
# process in slabs so that padded input + output fit comfortably
slab = 1 << 17
taps = torch.randn((slab, P), device=dev)
padded = torch.zeros((slab, 2 * P), device=dev)
padded[:, :P] = taps
n_slabs = -(-n_part // slab)


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


t_fft = timed(lambda: torch.fft.rfft(padded, dim=1))
t_pad = timed(lambda: torch.fft.rfft(torch.nn.functional.pad(taps, (0, P)), dim=1))
scale = n_part / slab
print(f"cuFFT R2C, padded input resident : {t_fft * scale:.2f} ms per step-equivalent ({t_fft:.3f} ms per {slab} transforms)")
print(f"cuFFT R2C incl. zero padding copy: {t_pad * scale:.2f} ms per step-equivalent")
print("k_ir_fft (bench.py, same batch)   : see roofline.kernel_ms.ir_fft (7.8 ms round 1)")
