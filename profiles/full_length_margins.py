import sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
from audio_mps_b200.model import _PsiLossFn
from audio_mps_b200 import HParams, PsiCMPS
from oracle.cmps_oracle import HP, damped_sine, random_raw_params
from tests.golden_util import load
from tests.util import rel, relc, set_raw
cuda = torch.device("cuda", 0)
for name, scan in [("psi_c1_full", False), ("psi_c4_d64_full_length", False), ("psi_c4_d64_full_length", True), ("psi_c3_d128_full_length", False)]:
    g = load(name)
    D, B, T, seed = int(g["D"]), int(g["B"]), int(g["T"]), int(g["seed"])
    hp = HP(bond_dim=D, minibatch_size=B)
    raw = random_raw_params(hp, np.random.default_rng(seed))
    data = damped_sine(B, T, hp.delta_t, np.random.default_rng(seed + 1))
    php = HParams(minibatch_size=B, bond_dim=D, delta_t=hp.delta_t, sigma=hp.sigma, h_reg=hp.h_reg, r_reg=hp.r_reg, initial_rank=None, A=hp.A, learning_rate=0.001)
    m = PsiCMPS(php, device=cuda); set_raw(m, raw)
    R = torch.view_as_real(m.R.detach()).clone().requires_grad_(); f = m.freqs.detach().clone().requires_grad_()
    p0 = torch.view_as_real(m.psi_0.detach()).clone().requires_grad_(); A = m.A.detach().clone().requires_grad_()
    x = torch.as_tensor(data, device=cuda)
    lpc = _PsiLossFn.apply(R, f, p0, A, x, m, scan)
    gR, gf, gp, gA = torch.autograd.grad(lpc.mean(), [R, f, p0, A])
    print(name, "scan" if scan else "seq", "loss %.1e" % rel(lpc.detach().cpu().numpy(), g["loss_f64"]),
          "gR %.1e" % relc(torch.view_as_complex(gR).cpu().numpy(), g["geff_R"]), "gf %.1e" % rel(gf.cpu().numpy(), g["geff_f"]),
          "gpsi %.1e" % relc(torch.view_as_complex(gp).cpu().numpy(), g["geff_psi0"]), "gA %.1e" % rel(gA.cpu().numpy(), g["geff_A"]))
