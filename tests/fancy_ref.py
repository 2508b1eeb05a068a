"""libjpeg's fancy (triangle-filter) chroma up-sampling restated in numpy float32 on the ORACLE's planes - the reference the
compose path (GPU: compose_colour_kernel, CPU simulation: sim_compose) is checked against.  Test infrastructure."""
import numpy as np

import oracle_ffi as O
from jpeg_rust_b200 import LAYOUT_SPEC, parse_descriptor


def fancy_reference(data, ext=0):
    """libjpeg's fancy (triangle-filter) chroma up-sampling restated in numpy float32 on the ORACLE's planes: the oracle
    (SPEC layout) replicates sub-sampled chroma, so every fx-th / fy-th sample of its full-size plane is the decoded
    chroma sample itself.  Same weights and operation order as compose_colour_kernel; colour conversion and truncation
    as decoder.rs:382-402."""
    o = O.decode(data, layout=O.LAYOUT_SPEC, ext=ext)
    H, W = o.height, o.width
    st, d, _buf = parse_descriptor(data, ext, LAYOUT_SPEC)
    hmax = max(d.comp[c].h for c in range(d.ncomp))
    vmax = max(d.comp[c].v for c in range(d.ncomp))
    planes = []
    for c in range(d.ncomp):
        full = o.planes[c].reshape(H, W).astype(np.float32) + np.float32(128.0 if c == 0 else 0.0)
        fx, fy = hmax // d.comp[c].h, vmax // d.comp[c].v
        if fx == 1 and fy == 1:
            planes.append(full)
            continue
        s = full[::fy, ::fx]                       # the component's own samples: (ceil(H/fy), ceil(W/fx))
        hc, wc = s.shape
        ys, xs = np.arange(H), np.arange(W)
        y0 = ys // fy if fy == 2 else ys
        x0 = xs // fx if fx == 2 else xs
        yn = np.clip(y0 + np.where(ys & 1, 1, -1), 0, hc - 1) if fy == 2 else y0
        xn = np.clip(x0 + np.where(xs & 1, 1, -1), 0, wc - 1) if fx == 2 else x0
        a, cc = s[np.ix_(y0, x0)], s[np.ix_(y0, xn)]
        if fy == 2:
            a = np.float32(0.75) * a + np.float32(0.25) * s[np.ix_(yn, x0)]
            cc = np.float32(0.75) * cc + np.float32(0.25) * s[np.ix_(yn, xn)]
        planes.append(np.float32(0.75) * a + np.float32(0.25) * cc if fx == 2 else a)
    if len(planes) == 1:
        u = np.clip(np.trunc(planes[0]), 0, 255).astype(np.uint8)
        return np.stack([u, u, u], axis=-1)
    y, cb, cr = planes
    f32 = np.float32
    r = cr * f32(1.402) + y
    g = cb * f32(-0.34413629) + (cr * f32(-0.71413629) + y)
    b = cb * f32(1.772) + y
    return np.stack([np.clip(np.trunc(v), 0, 255).astype(np.uint8) for v in (r, g, b)], axis=-1)
