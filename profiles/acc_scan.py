import sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
from audio_mps_b200 import PsiCMPS
from oracle import cref
from oracle.cmps_oracle import PsiCMPSOracle, damped_sine, random_raw_params
from tests.util import hp_pair, set_raw
dev = torch.device("cuda", 0)
for D, B, T in ((8, 2, 64000), (16, 2, 64000), (32, 2, 64000), (64, 2, 64000)):
    ohp, php = hp_pair(bond_dim=D, minibatch_size=B)
    for seed in (0, 1):
        raw = random_raw_params(ohp, np.random.default_rng(seed))
        data = damped_sine(B, T, ohp.delta_t, np.random.default_rng(seed + 10))
        m = PsiCMPS(php, device=dev); set_raw(m, raw)
        o = PsiCMPSOracle(ohp, raw, mode="f32", requires_grad=False)
        R, f, p0, A = cref.effective_from_oracle(o)
        ref = cref.psi_loss(R, f, p0, A, ohp.sigma, ohp.delta_t, data, mode="f64")
        seq = m.loss_per_clip(data, time_parallel=False).detach().cpu().numpy().astype(np.float64)
        scan = m.loss_per_clip_scan(data).detach().cpu().numpy().astype(np.float64)
        print(D, seed, "ref", ref, "seq err", np.abs(seq-ref)/np.abs(ref), "scan err", np.abs(scan-ref)/np.abs(ref))
