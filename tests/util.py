"""Shared tolerances.  The log-mel is ill-conditioned near its 1e-5 floor: an fp32 FFT is only accurate
to ~eps * (largest bin of the frame), and log(x+1e-5)*0.2 turns an absolute error d on a small bin x
into 0.2*d/(x+1e-5).  Two correct fp32 implementations (torch.stft vs rfft-of-frames) already differ
by 6e-7*rowmax in the linear domain (measured, scratch notes in DESIGN.md), so closeness is judged in
the LINEAR mel domain relative to the loudest bin."""
import numpy as np

LIN_REL_TO_MAX = 4e-6      # |lin - ref| <= 4e-6 * max(lin_ref over the frame / clip)
LIN_REL = 3e-5             # + 3e-5 * lin_ref  (exp/log round trip, fp32)


def to_linear(logmel):
    return np.exp((np.asarray(logmel, np.float64) - 0.9) / 0.2)


def assert_logmel_close(y, ref, scale_max, what=""):
    """y, ref: log-mel values (any matching shape); scale_max broadcastable linear-domain maxima."""
    ly, lr = to_linear(y), to_linear(ref)
    tol = LIN_REL_TO_MAX * np.asarray(scale_max, np.float64) + LIN_REL * lr + 1e-9
    bad = np.abs(ly - lr) > tol
    assert not bad.any(), f"{what}: {int(bad.sum())}/{bad.size} log-mel values outside tolerance; " \
                          f"worst lin err {np.abs(ly - lr).max():.3e}, worst err/tol {(np.abs(ly - lr) / tol).max():.2f}"


def rel_rows(x, y):
    import torch
    x, y = torch.as_tensor(x).float().cpu(), torch.as_tensor(y).float().cpu()
    return float(((x - y).norm(dim=-1) / y.norm(dim=-1)).max())
