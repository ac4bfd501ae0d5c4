"""CPU tests of the host side: hparams, raw->effective parameter chain, data formats, shard logic,
and that the C-ABI library loads and exports every symbol include/audiomps.h declares
(no compute calls without a GPU)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from audio_mps_b200 import HParams, PsiCMPS, RhoCMPS, _lib, default_hparams
from audio_mps_b200 import data as D
from audio_mps_b200.train import regulariser, shard_bounds
from oracle.cmps_oracle import HP, PsiCMPSOracle, RhoCMPSOracle, damped_sine, random_raw_params, total_loss
from tests.util import hp_pair, rel, relc, set_raw

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_abi_exports_every_declared_symbol(lib):
    hdr = open(os.path.join(ROOT, "include", "audiomps.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(amps_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(raw, name), f"{name} declared in audiomps.h but not exported"
        assert name in _lib.SYMBOLS, f"{name} has no ctypes binding"
    assert set(_lib.SYMBOLS) == declared
    assert lib.amps_version() == 200
    assert lib.amps_psi_grad_count(32) == 2 * 32 * 32 + 3 * 32 + 2
    assert lib.amps_psi_workspace_bytes(32, 64, 64000, 1) > 64 * 64000 * 32 * 8
    assert lib.amps_psi_workspace_bytes(200, 1, 10, 0) == 0      # unsupported D reports 0


@pytest.mark.parametrize("dt", [1 / 16000, 1 / 8000, 1 / 44100, 1e-3, 2.0 ** -14, 7.3e-5,
                                float(np.float32(0x800008 * 2.0 ** -37)), float(np.float32(0xC00400 * 2.0 ** -37))])
def test_time_table_is_the_float32_running_sum(lib, dt):
    """t_{k+1} = fl32(t_k + fl32(delta_t)) (model.py:16,281): the library's piecewise-exact generator
    (host twin of prep_ttab_kernel) against the sequential float32 sum, bit for bit."""
    n = 70000
    out = np.empty(n, np.float32)
    assert lib.amps_time_table_host(ctypes.c_double(dt), n, out.ctypes.data_as(ctypes.c_void_p)) == 0
    d, t = np.float32(dt), np.float32(0)
    ref = np.empty(n, np.float32)
    for k in range(n):
        ref[k] = t
        t = np.float32(t + d)
    assert np.array_equal(out.view(np.uint32), ref.view(np.uint32))


def test_no_cpu_fallback():
    hp = default_hparams()
    m = PsiCMPS(hp, device="cpu")
    with pytest.raises(RuntimeError):
        m.loss_per_clip(np.zeros((2, 16), np.float32))
    with pytest.raises(RuntimeError):
        m.sample(2, 8)
    with pytest.raises(RuntimeError):
        RhoCMPS(hp, device="cpu").loss_per_clip(np.zeros((2, 16), np.float32))


def test_hparams_parse():
    hp = default_hparams()
    hp.parse("bond_dim=32,minibatch_size=64,sigma=0.5,initial_rank=3")
    assert (hp.bond_dim, hp.minibatch_size, hp.sigma, hp.initial_rank) == (32, 64, 0.5, 3)
    assert isinstance(hp.bond_dim, int) and isinstance(hp.sigma, float)
    with pytest.raises(ValueError):
        hp.parse("nonexistent=1")
    assert hp.delta_t == 1 / 16000 and hp.A == 100.


@pytest.mark.parametrize("use_in", [False, True])
def test_effective_parameters_match_oracle(use_in):
    """model.py:31-52, 221-222: scale, the diagonal broadcast quirk, psi_0 normalisation."""
    ohp, php = hp_pair(bond_dim=5)
    rng = np.random.default_rng(0)
    raw = random_raw_params(ohp, rng)
    if use_in:
        R_in = (rng.standard_normal((5, 5)) + 1j * rng.standard_normal((5, 5))).astype(np.complex64)
        f_in = rng.standard_normal(5).astype(np.float32)
        o = PsiCMPSOracle(ohp, raw, R_in=R_in, freqs_in=f_in)
        m = PsiCMPS(php, R_in=R_in, freqs_in=f_in, device="cpu")
        set_raw(m, {"psi_x": raw["psi_x"], "psi_y": raw["psi_y"]})
    else:
        o = PsiCMPSOracle(ohp, raw)
        m = PsiCMPS(php, device="cpu")
        set_raw(m, raw)
    assert relc(m.R.detach().numpy(), o.R.detach().numpy()) <= 1e-6
    assert rel(m.freqs.detach().numpy(), o.freqs.detach().numpy()) <= 1e-6
    assert relc(m.psi_0.detach().numpy(), o.psi_0.detach().numpy()) <= 1e-6
    assert np.abs(np.diag(m.R.detach().numpy())).max() == 0           # tests/test_model.py:19-25
    # regulariser of train.py:55-60 and its gradient through the raw->effective chain
    reg = regulariser(m)
    oreg = ohp.h_reg * torch.sum(o.freqs ** 2) + ohp.r_reg * torch.sum(torch.conj(o.R) * o.R).real
    assert rel(float(reg), float(oreg)) <= 1e-5
    g = torch.autograd.grad(reg, [m.Rx, m.Ry, m.freqs_raw])
    og = torch.autograd.grad(oreg, [o.vars["Rx"], o.vars["Ry"], o.vars["freqs"]])
    for a, b in zip(g, og):
        assert rel(a.numpy(), b.numpy()) <= 1e-5


def test_rho0_matches_oracle():
    ohp, php = hp_pair(bond_dim=4, initial_rank=3)
    raw = random_raw_params(ohp, np.random.default_rng(1), rho=True)
    assert raw["Wx"].shape == (3, 4)
    o = RhoCMPSOracle(ohp, raw)
    m = RhoCMPS(php, device="cpu")
    set_raw(m, raw)
    assert relc(m.rho_0.detach().numpy(), o.rho_0.detach().numpy()) <= 1e-6
    r = m.rho_0.detach().numpy()
    assert abs(np.trace(r) - 1) < 1e-6 and np.abs(r - r.conj().T).max() < 1e-6   # tests/test_model.py:41-48


def test_single_step_primitives_trivial_update():
    """tests/test_model.py:69-83, 124-138 (H = R = 0 leaves the state untouched)."""
    _, php = hp_pair(bond_dim=7, h_reg=2 / (np.pi * 16000) ** 2, r_reg=2 / (np.pi * 16000))
    z = np.zeros((7, 7), np.complex64)
    f0 = np.zeros(7, np.float32)
    sig = np.random.default_rng(0).random(8).astype(np.float32)
    m = PsiCMPS(php, freqs_in=f0, R_in=z, device="cpu")
    st = m.psi_0.unsqueeze(0).repeat(8, 1)
    np.testing.assert_allclose(m._update_ancilla_psi(st, sig, 0.).detach().numpy(), st.detach().numpy(), rtol=1e-6)
    r = RhoCMPS(php, freqs_in=f0, R_in=z, device="cpu")
    sr = r.rho_0.unsqueeze(0).repeat(8, 1, 1)
    np.testing.assert_allclose(r._update_ancilla_rho(sr, sig, 0.).detach().numpy(), sr.detach().numpy(), rtol=1e-6)


def test_single_step_primitives_match_oracle():
    ohp, php = hp_pair(bond_dim=6, sigma=0.3, A=2.0)
    raw = random_raw_params(ohp, np.random.default_rng(2))
    o = PsiCMPSOracle(ohp, raw)
    m = PsiCMPS(php, device="cpu")
    set_raw(m, raw)
    psi = o.psi_0.unsqueeze(0).repeat(3, 1)
    sig = np.array([0.1, -0.2, 0.05], np.float32)
    a = m._update_ancilla_psi(psi.detach(), sig, 0.37).detach().numpy()
    b = o._update_ancilla_psi(psi, torch.tensor(sig), np.float32(0.37)).detach().numpy()
    assert relc(a, b) <= 1e-5
    assert rel(m._expectation(psi.detach(), 0.37).detach().numpy(),
               o._expectation(psi, np.float32(0.37)).detach().numpy()) <= 1e-5


def test_damped_sine_matches_oracle_generator():
    a = D.damped_sine(4, 300, 1 / 16000, np.random.default_rng(5))
    b = damped_sine(4, 300, 1 / 16000, np.random.default_rng(5))
    assert a.shape == (4, 300) and a.dtype == np.float32            # tests/test_data.py:12-16
    np.testing.assert_array_equal(a, b)
    assert np.abs(a).max() <= 1.0


def test_tfrecord_roundtrip(tmp_path):
    """{audio: float32[sample_duration]} TFRecords (data.py:27-43) without TensorFlow."""
    clips = np.random.default_rng(0).standard_normal((5, 64)).astype(np.float32)
    path = str(tmp_path / "guitar.tfrecords")
    D.write_tfrecords(path, clips)
    recs = list(D.iter_tfrecords(path))
    assert len(recs) == 5
    np.testing.assert_array_equal(D.parse_example_float_feature(recs[3], "audio"), clips[3])
    with pytest.raises(KeyError):
        D.parse_example_float_feature(recs[0], "pitch")
    hp = HParams(minibatch_size=2, delta_t=1 / 16000)
    it = D.get_audio(str(tmp_path), "guitar", hp, sample_duration=64)
    b0, b1, b2, b3 = next(it), next(it), next(it), next(it)
    assert b0.shape == (2, 64) and b2.shape == (1, 64)              # ragged last batch, then repeat
    np.testing.assert_array_equal(b3, clips[:2])
    with pytest.raises(ValueError):
        next(D.get_audio(str(tmp_path), "guitar", hp, sample_duration=32))


def test_shard_bounds():
    for gb, world in ((2048, 8), (64, 1), (10, 4), (3, 4)):
        spans = [shard_bounds(gb, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == gb
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1


def test_tf_checkpoint_round_trip(tmp_path):
    """TensorFlow V2 checkpoint (tensor bundle) writer/reader: table structure, checksums, names."""
    import struct
    from audio_mps_b200 import tf_checkpoint as tfc
    assert tfc.crc32c(b"123456789") == 0xE3069283                    # CRC-32C check value
    rng = np.random.default_rng(0)
    tensors = {"model/Rx": rng.standard_normal((7, 7)).astype(np.float32),
               "model/Ry": rng.standard_normal((7, 7)).astype(np.float32),
               "model/freqs": rng.standard_normal(7).astype(np.float32),
               "model/psi_x": rng.standard_normal(7).astype(np.float32),
               "model/psi_y": rng.standard_normal(7).astype(np.float32),
               "model/A": np.float32(100.0), "global_step": np.int64(1617),
               "model/Rx/Adam": np.zeros((7, 7), np.float32), "beta1_power": np.float32(0.9)}
    for i in range(40):                                              # several restart intervals
        tensors[f"extra/v{i:02d}"] = rng.standard_normal(3).astype(np.float32)
    prefix = str(tmp_path / "model.ckpt-1617")
    tfc.write_tf_checkpoint(prefix, tensors)
    idx = open(prefix + ".index", "rb").read()
    assert struct.unpack("<Q", idx[-8:])[0] == 0xDB4775248B80FB57    # LevelDB table magic
    back = tfc.read_tf_checkpoint(prefix)
    assert set(back) == set(tensors)
    for k, v in tensors.items():
        assert back[k].dtype == np.asarray(v).dtype and np.array_equal(back[k], np.asarray(v)), k
    assert tfc.latest_checkpoint(str(tmp_path)) == prefix
    # corruption is detected
    bad = bytearray(open(prefix + ".data-00000-of-00001", "rb").read())
    bad[5] ^= 0xFF
    open(prefix + ".data-00000-of-00001", "wb").write(bytes(bad))
    with pytest.raises(ValueError):
        tfc.read_tf_checkpoint(prefix)


def test_tf_checkpoint_state_file_of_the_reference(tmp_path):
    """`checkpoint` state files as TF writes them (the layout of /root/reference/logging/checkpoint)."""
    from audio_mps_b200 import tf_checkpoint as tfc
    (tmp_path / "checkpoint").write_text('model_checkpoint_path: "model.ckpt-1617"\n'
                                         'all_model_checkpoint_paths: "model.ckpt-1436"\n'
                                         'all_model_checkpoint_paths: "model.ckpt-1617"\n')
    assert tfc.latest_checkpoint(str(tmp_path)) == str(tmp_path / "model.ckpt-1617")
    assert tfc.latest_checkpoint(str(tmp_path / "nothing")) is None


def test_model_tf_checkpoint_names(tmp_path):
    """Psi and Rho models save / load under the reference's variable names (model.py:19,32-50,122-126,
    215-219; scope "model", train.py:49)."""
    from audio_mps_b200 import HParams, PsiCMPS, RhoCMPS, tf_checkpoint as tfc
    hp = HParams(minibatch_size=2, bond_dim=5, delta_t=1 / 16000, sigma=1e-4, h_reg=1e-9, r_reg=0.1,
                 initial_rank=3, A=100., learning_rate=1e-3)
    for cls, names in ((PsiCMPS, {"A", "Rx", "Ry", "freqs", "psi_x", "psi_y"}),
                       (RhoCMPS, {"A", "Rx", "Ry", "freqs", "Wx", "Wy"})):
        m = cls(hp, device="cpu", seed=1)
        prefix = str(tmp_path / cls.__name__ / "model.ckpt-7")
        m.save_tf_checkpoint(prefix, global_step=7)
        got = tfc.read_tf_checkpoint(prefix)
        assert set(got) == {"model/" + n for n in names} | {"global_step"}
        m2 = cls(hp, device="cpu", seed=2)
        assert m2.load_tf_checkpoint(str(tmp_path / cls.__name__)) == 7     # via the state file
        for (n, a), (_, b) in zip(m.named_parameters(), m2.named_parameters()):
            assert torch.equal(a, b), n
    with pytest.raises(ValueError):                                  # shape mismatch is an error
        PsiCMPS(HParams(**{**hp.values(), "bond_dim": 6}), device="cpu").load_tf_checkpoint(
            str(tmp_path / "PsiCMPS" / "model.ckpt-7"))


def test_tf_checkpoint_reads_snappy_compressed_index(tmp_path):
    """Index tables written with snappy block compression (type byte 1) are decoded too: rewrite a
    checkpoint's index with its data block stored as snappy (literals plus one back-reference)."""
    import struct
    from audio_mps_b200 import tf_checkpoint as tfc
    # the decompressor itself: literal "abcd", copy(offset 4, len 8) -> "abcdabcdabcd", long literal
    raw = b"abcd" * 3 + bytes(range(70))
    comp = tfc._put_varint(len(raw)) + bytes([(4 - 1) << 2]) + b"abcd" + bytes([((8 - 4) << 2) | 1, 4]) \
        + bytes([60 << 2, 70 - 1]) + bytes(range(70))
    assert tfc._snappy_decompress(comp) == raw
    # a whole index file with a compressed data block
    tensors = {"model/A": np.float32(3.5), "model/freqs": np.arange(5, dtype=np.float32)}
    prefix = str(tmp_path / "m.ckpt-1")
    tfc.write_tf_checkpoint(prefix, tensors)
    idx = open(prefix + ".index", "rb").read()
    footer = idx[-48:]
    _, p = tfc._get_varint(footer, 0)
    _, p = tfc._get_varint(footer, p)
    ioff, p = tfc._get_varint(footer, p)
    isz, p = tfc._get_varint(footer, p)
    (_, handle), = list(tfc._block_entries(tfc._read_block(idx, ioff, isz)))
    boff, q = tfc._get_varint(handle, 0)
    bsz, q = tfc._get_varint(handle, q)
    block = idx[boff:boff + bsz]
    lit = b""                                      # literal-only snappy stream, 60-byte pieces
    for s0 in range(0, len(block), 60):
        piece = block[s0:s0 + 60]
        lit += bytes([(len(piece) - 1) << 2]) + piece
    cblock = tfc._put_varint(len(block)) + lit
    f = bytearray()
    f += cblock + b"\x01" + struct.pack("<I", tfc.masked_crc32c(cblock + b"\x01"))
    dh = tfc._put_varint(0) + tfc._put_varint(len(cblock))
    mh = tfc._emit_block(f, tfc._build_block([]))
    ih = tfc._emit_block(f, tfc._build_block([(b"\xff", dh)], restart_interval=1))
    foot = mh + ih
    f += foot + b"\x00" * (40 - len(foot)) + struct.pack("<Q", 0xDB4775248B80FB57)
    open(prefix + ".index", "wb").write(bytes(f))
    back = tfc.read_tf_checkpoint(prefix)
    assert float(back["model/A"]) == 3.5 and np.array_equal(back["model/freqs"], tensors["model/freqs"])


def test_host_generators_match_the_oracle_generators():
    """audio_mps_b200.data.random_raw_params / sample_noise (what bench.py feeds the CUDA path) draw the
    same numbers as the oracle-side generators the goldens were minted with."""
    from audio_mps_b200.data import random_raw_params as prp, sample_noise as psn
    from oracle.mint_golden_r2 import sample_noise as osn
    hp = HP(bond_dim=16)
    a, b = prp(16, hp.A, np.random.default_rng(3)), random_raw_params(hp, np.random.default_rng(3))
    assert set(a) == set(b) and all(np.array_equal(a[k], b[k]) for k in b)
    assert np.array_equal(psn(hp.sigma, hp.delta_t, 100, 3, 5), osn(hp, 100, 3, 5))


def test_waveform_image_and_summaries(tmp_path):
    """utils.waveform_plot / the summaries block of train.py:62-85 without TF, tfplot or matplotlib."""
    import types

    import torch
    from audio_mps_b200.train_cli import waveform_image, write_summaries
    from tensorboard.backend.event_processing.event_accumulator import EventAccumulator
    from torch.utils.tensorboard import SummaryWriter
    x = np.sin(np.linspace(0, 20, 4000)).astype(np.float32)
    img = waveform_image(x)
    assert img.shape == (3, 300, 300) and img.dtype == np.uint8
    inner = img[0, 13:-13, 13:-13]
    assert (inner == 40).any(axis=0).all()                     # every pixel column carries part of the curve
    assert (img[0, 12, 12:-12] == 0).all() and (img[0, 0] == 255).all()      # axes box, white margin
    assert waveform_image(np.zeros(10)).shape == (3, 300, 300)  # constant signal: no division by zero
    assert waveform_image(np.zeros(0)).shape == (3, 300, 300)
    tb = SummaryWriter(str(tmp_path))
    model = types.SimpleNamespace(freqs=torch.linspace(-3000.0, 3000.0, 8))
    batch = np.stack([x, -x, 0.5 * x, x, x, x, x])
    write_summaries(tb, 7, batch, model, 16000, visualize=True, samples=batch[:2])
    tb.close()
    ea = EventAccumulator(str(tmp_path), size_guidance={"audio": 0, "images": 0, "histograms": 0})
    ea.Reload()
    tags = ea.Tags()
    assert sorted(tags["audio"]) == [f"data/{i}" for i in range(5)]          # max_outputs = 5
    assert tags["histograms"] == ["frequencies"]
    assert len([t for t in tags["images"] if t.startswith("data_waveform/")]) == 7
    assert len([t for t in tags["images"] if t.startswith("sample_waveform/")]) == 2
    assert ea.Audio("data/0")[0].sample_rate == 16000 and ea.Audio("data/0")[0].step == 7
