"""Extracts the lower envelope of the log-posterior trace the reference publishes for its README
kernel_adapt() run (man/figures/get_-1.png: `plot(get_logpost(), type = "l")` right after
README.md:250-269) into tests/golden/readme_am_logpost_envelope.npz.

Run here (the reference tree is not on the GPU box):  python tests/golden/make_readme_am_envelope.py
The figure is a 672 x 480 base-graphics plot; the plot box spans columns 78..631 / rows 78..382, the x
ticks 0..5000 sit at columns 99..611 (102.4 px per 1000 iterations) and the y ticks -2830 / -2845 at rows
142 / 373 (15.4 px per unit) - measured from the tick marks by this script, asserted below."""
import os
import sys

import numpy as np
from PIL import Image

REF = "/root/reference/man/figures/get_-1.png"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "readme_am_logpost_envelope.npz")


def main():
    im = np.array(Image.open(REF).convert("L")).astype(int)
    H, W = im.shape
    dark = im < 128
    yt = [i for i in range(H) if dark[i, 70:78].sum() >= 6]
    xt = [j for j in range(W) if dark[383:391, j].sum() >= 6]
    assert yt == [142, 219, 296, 373] and xt == [99, 201, 303, 406, 508, 611], (yt, xt)
    dark = im < 160
    cols = np.arange(79, 631)
    lo = np.full(cols.size, -1, dtype=np.int32)      # lowest dark pixel of the curve in each column
    hi = np.full(cols.size, -1, dtype=np.int32)
    for n, j in enumerate(cols):
        ys = np.where(dark[80:380, j])[0]
        if ys.size:
            lo[n], hi[n] = ys.max() + 80, ys.min() + 80
    np.savez_compressed(OUT, cols=cols, lo=lo, hi=hi, x0_col=99.0, px_per_iter=0.1024,
                        y_ref=-2830.0, y_ref_row=142.0, px_per_unit=15.4)
    print("wrote", OUT, lo[20:40])


if __name__ == "__main__":
    sys.exit(main())
