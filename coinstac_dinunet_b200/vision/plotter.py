"""Progress plots: one PNG per header group of ``log_header`` (``'Loss|Accuracy,F1'``).

Parity: coinstac_dinunet/vision/plotter.py:20-65 (file name ``<key>_<plot_id>.png``, raw curve
plus rolling mean).  matplotlib is imported lazily; when it is not installed (as on the B200
image) the same series are written as ``<key>_<plot_id>.csv`` next to where the PNG would go,
so training never fails because of a plotting dependency.
"""
import math as _math
import os as _os

import numpy as _np

COLORS = ['blue', 'maroon', 'magenta', 'teal', 'red', 'blueviolet', 'brown', 'cadetblue',
          'chartreuse', 'coral', 'darkslateblue', 'cornflowerblue', 'indigo', 'black', 'cyan', 'navy']

_plt = None


def _pyplot():
    global _plt
    if _plt is None:
        try:
            import matplotlib
            matplotlib.use('agg')
            import matplotlib.pyplot as plt
            if not hasattr(plt, 'subplots'):      # a stub module (e.g. installed by a harness): treat as absent
                raise ImportError('matplotlib.pyplot has no subplots')
            plt.rcParams['figure.figsize'] = [16, 9]
            _plt = plt
        except Exception:
            _plt = False
    return _plt


def _rolling_mean(a, window):
    """Trailing mean with ``min_periods=1`` (pandas ``rolling(window).mean()`` semantics)."""
    a = _np.asarray(a, dtype=_np.float64)
    csum = _np.cumsum(_np.vstack([_np.zeros((1, a.shape[1])), a]), axis=0)
    idx = _np.arange(1, a.shape[0] + 1)
    lo = _np.maximum(idx - window, 0)
    return (csum[idx] - csum[lo]) / (idx - lo)[:, None]


def plot_progress(cache, log_dir, plot_keys=(), num_points=15, epoch=None):
    header_spec = cache.get('log_header')
    for key in plot_keys:
        rows = cache.get(key, [])
        if header_spec is None or len(rows) == 0:
            continue
        table = _np.asarray(rows, dtype=_np.float64)
        if table.ndim != 2:
            continue
        col = 0
        for plot_id, group in enumerate(str(header_spec).split('|')):
            names = group.split(',')
            lo, hi = col, col + len(names)
            block = table[:, lo:hi]
            if block.shape[1] != len(names) or _np.sum(block) <= 0:
                continue
            col = hi
            window = max(block.shape[0] // num_points, 3)
            smooth = _rolling_mean(block, window)
            target = _os.path.join(log_dir, f'{key}_{plot_id}')
            plt = _pyplot()
            if not plt:
                with open(target + '.csv', 'w') as fp:
                    fp.write(','.join(names + [f'{n}_rolling' for n in names]) + '\n')
                    for raw, sm in zip(block, smooth):
                        fp.write(','.join(f'{v:.6g}' for v in [*raw, *sm]) + '\n')
                continue
            plt.clf()
            fig, ax = plt.subplots()
            xs = _np.arange(block.shape[0])
            for j, name in enumerate(names):
                c = COLORS[(lo + j) % len(COLORS)]
                ax.plot(xs, block[:, j], alpha=0.11, color=c)
                ax.plot(xs, smooth[:, j], color=c, label=name)
            ax.set_title(str(key).upper())
            ax.legend()
            if epoch and epoch != block.shape[0] and block.shape[0] // epoch > 0:
                ticks = list(range(0, block.shape[0], block.shape[0] // epoch)) + [block.shape[0] - 1]
                step = int(_math.log(len(ticks) + 1) + len(ticks) // num_points + 1)
                ax.set_xticks(ticks[::step])
                ax.set_xticklabels(list(range(len(ticks)))[::step])
            ax.set_xlabel('Epochs')
            fig.savefig(target + '.png')
            plt.close('all')
