"""Distribution of the FP64 reciprocal-square-root seed's residual e = 1 - x*y0^2 on the device (MUFU.RSQ64H on the
high word, csrc/pair_math.cuh rsqrt_masked<RSQ_SEED>), and the relative error of the refined values per digits class.
The one-step refinement y0*(1 + e/2 + delta) has error 3/8 e^2 - delta: delta is chosen from mean(e^2) measured here.

    python tools/seed_error.py            # prints a markdown table
"""
import sys
import os

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import sctl_b200 as sb  # noqa: E402


def main():
    sb.load_library()
    n = 1 << 24
    g = torch.Generator(device="cuda").manual_seed(7)
    rows = []
    print("| inputs | mean e | mean e^2 | min e | max e | 3/8 mean e^2 | 3/8 max e^2 |")
    print("|---|---|---|---|---|---|---|")
    for label, lo, hi in (("x uniform in [1, 4)", 1.0, 4.0), ("x uniform in [1e-8, 3]", 1e-8, 3.0), ("x = r^2, r uniform in (0, 1.7)", None, None),
                          ("x log-uniform in [1e-300, 1e300]", "log", None)):
        u = torch.rand(n, dtype=torch.float64, device="cuda", generator=g)
        if lo is None:
            x = (u * 1.7) ** 2 + 1e-30
        elif lo == "log":
            x = torch.pow(torch.tensor(10.0, dtype=torch.float64, device="cuda"), (u - 0.5) * 600.0)
        else:
            x = lo + (hi - lo) * u
        y0 = torch.empty_like(x)
        sb.rsqrt_device(x, y0, digits=0)  # seed only
        torch.cuda.synchronize()
        xs, ys = x.cpu().numpy().astype(np.longdouble), y0.cpu().numpy().astype(np.longdouble)
        e = 1.0 - xs * ys * ys
        rows.append(e)
        print("| %s | %.3e | %.3e | %.3e | %.3e | %.3e | %.3e |" % (label, e.mean(), (e * e).mean(), e.min(), e.max(), 0.375 * (e * e).mean(), 0.375 * (e * e).max()))
    print()
    print("| digits | max rel err of rsqrt | mean rel err | (x = r^2, r uniform) |")
    print("|---|---|---|---|")
    u = torch.rand(n, dtype=torch.float64, device="cuda", generator=g)
    x = (u * 1.7) ** 2 + 1e-30
    xs = x.cpu().numpy().astype(np.longdouble)
    exact = 1.0 / np.sqrt(xs)
    for digits in (0, 5, 6, 8, 9, 11, 12, 13, 15, -1):
        y = torch.empty_like(x)
        sb.rsqrt_device(x, y, digits=digits)
        torch.cuda.synchronize()
        rel = (y.cpu().numpy().astype(np.longdouble) - exact) / exact
        print("| %d | %.3e | %.3e | |" % (digits, np.abs(rel).max(), rel.mean()))


if __name__ == "__main__":
    main()
