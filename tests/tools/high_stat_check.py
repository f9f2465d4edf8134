"""High-statistics distribution check: per-bin mean / variance of 2.6e8 traces per configuration
against the closed form (oracle.closed_form_moments).  Statistical error on a bin mean is < 1e-3 LSB,
so systematic deviations of the GPU samplers at the 1e-3 LSB level would show.  GPU only; uses the
oracle as the checker (test infrastructure)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from digicamtoy_b200 import NTraceGenerator, workloads
from digicamtoy_b200.stats import CubeStats
from oracle import reference_port as orc

def on_grid_moments(cfg, kw):
    """sub_binning > 0: V_b = g sum_m h(m dt) S_(b-m) with S the clusters of Poisson(rate dt) primaries
    on every simulated sample (reference :259-267, :306-312, :329-348); pe exist from simulated bin 0."""
    dt = float(kw['time_sampling'])
    n_drop, B = cfg.n_drop, len(cfg.sampling_bins) - cfg.n_drop
    n_taps = int(np.floor(cfg.template_x[-1] / dt + 1e-9)) + 1
    h = np.asarray(cfg.pulse_template(np.arange(n_taps) * dt), dtype=np.float64)
    xt, g, s1 = cfg.crosstalk[:, None], cfg.gain[:, None], cfg.sigma_1[:, None]
    lam = cfg.nsb_rate[:, None] * dt
    m1 = 1 / (1 - xt)
    m2 = xt / (1 - xt) ** 3 + m1 ** 2 + s1 ** 2 * m1
    s1h = np.array([h[:min(n_taps, b + n_drop + 1)].sum() for b in range(B)])[None, :]
    s2h = np.array([(h[:min(n_taps, b + n_drop + 1)] ** 2).sum() for b in range(B)])[None, :]
    return cfg.baseline[:, None] + g * lam * m1 * s1h, g ** 2 * lam * m2 * s2h + cfg.sigma_e[:, None] ** 2 + 1 / 12


def check(name, kw, n_events, chunk=16384, kernel='auto'):
    g = NTraceGenerator(seed=12345, saturate=(-2000, 6000), kernel=kernel, **kw)
    st = CubeStats(g)
    buf = torch.empty((chunk, g.params.n_pixels, g.params.n_bins), dtype=torch.int16, device='cuda')
    done = 0
    while done < n_events:
        n = min(chunk, n_events - done)
        g.generate(n, first_event=done, out=buf[:n])
        st.accumulate(buf[:n])
        done += n
    r = st.result()
    P, B = g.params.n_pixels, g.params.n_bins
    N = n_events * P
    m_hat = r['bin_sum'] / N
    v_hat = r['bin_sumsq'] / N - m_hat ** 2
    cfg = orc.make_config(**kw)
    mean, var = on_grid_moments(cfg, kw) if kw.get('sub_binning', 0) > 0 else orc.closed_form_moments(cfg, n_grid=40001)
    mean_b, var_b = mean.mean(axis=0), (var + mean ** 2).mean(axis=0) - mean.mean(axis=0) ** 2
    z = (m_hat - mean_b) / np.sqrt(var_b / N)
    print('%s [kernel %s]: %d events, %.2e traces per bin' % (name, g.kernel, n_events, N))
    print('   mean: max |dev| %.2e LSB, max |z| %.2f, mean z %.2f' % (np.abs(m_hat - mean_b).max(), np.abs(z).max(), z.mean()))
    print('   var : max rel dev %.2e (rounding term 1/12 included in the closed form)' % np.abs(v_hat / var_b - 1).max())
    return z

check('ac_dc NSB 1 GHz, no pulse', workloads.c2_ac_dc(1.0, 0.), 200000)                       # auto: tensor kernel
check('ac_dc NSB 1 GHz, no pulse', workloads.c2_ac_dc(1.0, 0.), 200000, kernel='sweep')
check('ac_dc NSB 0.66 GHz, no pulse', workloads.c2_ac_dc(0.66, 0.), 200000)                    # auto: tensor kernel
check('ac_dc NSB 0.04 GHz, no pulse', workloads.c2_ac_dc(0.04, 0.), 200000)
check('dark.yml, no pulse', workloads.c1_dark(0.), 100000)
check('ac_dc NSB 1 GHz on the sampling grid (sub_binning=1), no pulse', dict(sub_binning=1, **workloads.c2_ac_dc(1.0, 0.)), 200000)
check('ac_dc NSB 0.125 GHz on the sampling grid (sub_binning=1), no pulse', dict(sub_binning=1, **workloads.c2_ac_dc(0.125, 0.)), 200000)
