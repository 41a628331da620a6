"""CPU tests of the product's host-side mirror of the reference classes against the oracle: two
independent restatements (closed-form matrices vs the oracle's literal IFFT modem) must agree."""
import numpy as np
import pytest

import chest_b200
from chest_b200.modulation import FBMC, OFDM, SignalConstellation
from chest_b200.estimation import (ImaginaryInterferenceCancellationAtPilotPosition as IIC,
                                   PilotSymbolAidedChannelEstimation as PSACE)
from chest_b200.channel import FastFading
from chest_b200.simulation import DoublySelectiveSimulation


def test_constellation_tables_match_oracle():
    from oracle.signal_constellation import SignalConstellation as Ref
    for order, method in ((16, "PAM"), (4, "PAM"), (256, "QAM"), (16, "QAM"), (4, "QAM")):
        a, b = SignalConstellation(order, method), Ref(order, method)
        assert np.allclose(a.SymbolMapping, b.SymbolMapping, rtol=0, atol=1e-15)
        assert np.array_equal(a.BitMapping, b.BitMapping)
        bits = np.random.default_rng(0).integers(0, 2, 40 * a.BitMapping.shape[1])
        assert np.array_equal(a.Symbol2Bit(a.Bit2Symbol(bits)), bits)


def test_modem_matrices_match_oracle(ds_default):
    S = ds_default
    f = FBMC(24, 30, 15e3, 15e3 * 24, 0, False, "Hermite-OQAM", 8, 0, True)
    assert f.Nr["SamplesTotal"] == S["N"] == 540
    G = f.GetTXMatrix()
    assert np.max(np.abs(G - S["wf"]["F"]["G"])) < 1e-13
    assert np.max(np.abs(f.GetRXMatrix().conj().T - S["wf"]["F"]["Q"])) < 1e-13
    assert np.max(np.abs(f.GetFBMCMatrix() - S["D0"])) < 1e-13
    o = OFDM(24, 14, 15e3, 15e3 * 24, 0, False, 1 / 15e3 / 14, S["ofdm"].PHY["ZeroGuardTimeLength"])
    assert o.Implementation["CyclicPrefix"] == 2 and o.Implementation["ZeroGuardSamples"] == 88
    assert np.max(np.abs(o.GetTXMatrix() - S["wf"]["O"]["G"])) < 1e-13
    assert np.max(np.abs(o.GetRXMatrix().conj().T - S["wf"]["O"]["Q"])) < 1e-13


def test_modem_with_intermediate_frequency_matches_oracle():
    """SimpleVersion_DoublyFlat.m geometry: 12 subcarriers, IF = 20 subcarriers, fs = 2.52 MHz."""
    from oracle.fbmc import FBMC as RF
    from oracle.ofdm import OFDM as RO
    a = FBMC(12, 30, 15e3, 15e3 * 14 * 12, 15e3 * 20, False, "Hermite-OQAM", 8, 0, True)
    b = RF(12, 30, 15e3, 15e3 * 14 * 12, 15e3 * 20, False, "Hermite-OQAM", 8, 0, True)
    assert a.Nr["SamplesTotal"] == b.Nr["SamplesTotal"] == 3780
    assert np.max(np.abs(a.GetTXMatrix() - b.GetTXMatrix())) < 1e-12
    c = OFDM(12, 15, 15e3, 15e3 * 14 * 12, 15e3 * 20, False, 0, (8 - 1 / 2) * 1 / 15e3 * 1 / 2)
    d = RO(12, 15, 15e3, 15e3 * 14 * 12, 15e3 * 20, False, 0, (8 - 1 / 2) * 1 / 15e3 * 1 / 2)
    assert c.Nr["SamplesTotal"] == d.Nr["SamplesTotal"] == 3780
    assert np.max(np.abs(c.GetTXMatrix() - d.GetTXMatrix())) < 1e-12
    assert np.max(np.abs(c.GetRXMatrix() - d.GetRXMatrix())) < 1e-12


def test_precoders_match_oracle(ds_default):
    S = ds_default
    pm_o, pm_f, pm_aux = DoublySelectiveSimulation._pilot_matrices(24, 1)
    assert np.array_equal(pm_f, S["pm_f"]) and np.array_equal(pm_aux, S["pm_aux"]) and np.array_equal(pm_o, S["pm_o"])
    a = IIC("Auxiliary", pm_aux, S["D0"], 28, 4.685)
    assert (a.NrPilotSymbols, a.NrDataSymbols, a.NrAuxiliarySymbols) == (16, 640, 64)
    assert np.max(np.abs(a.PrecodingMatrix - S["aux"].PrecodingMatrix)) < 1e-12
    assert abs(a.DataPowerReduction - S["aux"].DataPowerReduction) < 1e-14
    c = IIC("Coding", pm_f, S["D0"], 20, 4)
    assert c.NrDataSymbols == 688
    assert np.max(np.abs(c.PrecodingMatrix - S["cod"].PrecodingMatrix)) < 1e-12
    assert np.allclose(c.SIR_dB, S["cod"].SIR_dB) and np.allclose(a.SIR_dB, S["aux"].SIR_dB)
    with pytest.raises(ValueError):
        IIC("Coding", pm_f, S["D0"], 200, 4)          # overlapping spreading sets (IIC.m:116-118)
    with pytest.raises(ValueError):
        IIC("Nonsense", pm_f, S["D0"], 20, 4)


def test_pilot_pattern_and_interpolators():
    d = PSACE("Diamond", [[12, 6], [30, 8]], "linear")
    assert d.NrPilotSymbols == 8                       # SV.m:57-66
    assert d.GetAuxiliaryMatrix(1).min() == -1 and np.sum(d.GetAuxiliaryMatrix(1) == -1) == 8
    r = PSACE("Rectangular", [[12, 6], [30, 8]], "FullAverage")
    assert np.allclose(r.ChannelInterpolation(np.full(r.NrPilotSymbols, 2 + 1j)), 2 + 1j)
    mv = PSACE("Diamond", [[12, 6], [30, 8]], "MovingBlockAverage", [12, 30])
    M = mv.GetInterpolationMatrix()
    assert np.allclose(M.sum(axis=1), 1)
    lin = d.GetInterpolationMatrix()
    assert lin.shape == (360, 8) and np.allclose(lin.sum(axis=1), 1)
    with pytest.raises(NotImplementedError):
        PSACE("Diamond", [[12, 6], [30, 8]], "MMSE")   # PSACE.m:110-111: the reference errors too
    with pytest.raises(ValueError):
        PSACE("Hexagonal", [[12, 6], [30, 8]], "linear")


def test_fast_fading_constructor_tables():
    ch = FastFading(15e3 * 24, "VehicularA", 540, 1158.18, "Jakes", 200, 1, 1, False, create_device=False)
    assert np.allclose(ch.Implementation["PowerDelayProfileNormalized"], [0.97981283, 0.02018717])
    ch2 = FastFading(15e3 * 14 * 14, "VehicularA", 7350, 1158.18, "Jakes", 200, 1, 1, False, create_device=False)
    assert list(ch2.Implementation["IndexDelayTaps"]) == [0, 1, 2, 3, 5, 7]
    assert np.allclose(ch2.Implementation["PowerDelayProfileNormalized"][[0, 1, 2, 3, 5, 7]],
                       [0.485, 0.3853, 0.0611, 0.0485, 0.0153, 0.0049], atol=5e-5)
    from oracle.fast_fading import FastFading as Ref
    for name in ("PedestrianA", "VehicularB", "TDL-A_30ns", "ExtendedPedestrianA"):
        a = FastFading(7.68e6, name, 100, 100.0, "Jakes", 10, create_device=False)
        b = Ref(7.68e6, name, 100, 100.0, "Jakes", 10)
        assert np.allclose(a.Implementation["PowerDelayProfileNormalized"], b.Implementation["PowerDelayProfileNormalized"])
    with pytest.raises(ValueError):
        FastFading(1e6, "NoSuchModel", 10, 10.0, "Jakes", 4, create_device=False)
    d = FastFading(360e3, "VehicularA", 540, 1158.18, "Discrete-Jakes", 4, 2, 2, create_device=False)   # FF.m:151-175
    r = Ref(360e3, "VehicularA", 540, 1158.18, "Discrete-Jakes", 4)
    assert np.allclose(d.Implementation["DiscreteDopplerSpectrum"], r.Implementation["DiscreteDopplerSpectrum"], atol=1e-15)
    assert d.Nr["txAntennas"] == 2 and d.Nr["rxAntennas"] == 2
    slow = FastFading(360e3, "VehicularA", 540, 100.0, "Discrete-Uniform", 4, create_device=False)      # FF.m:146-149
    assert slow.PHY["MaximumDopplerShift"] == 0
    with pytest.raises(ValueError):
        FastFading(1e6, "Flat", 10, 10.0, "Gaussian", 4, create_device=False)


def test_product_fails_loudly_without_gpu():
    """No CPU fallback: on a box without an sm_100 device the product refuses to run."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(chest_b200.ChestError):
        chest_b200.DeviceContext(0)
    with pytest.raises(chest_b200.ChestError):
        FBMC().Modulation(np.zeros((12, 30)))


def test_iic_tie_group_rule(ds_default):
    """IIC.m:72-73,113-114 at the default geometry: the (N+1)-th largest interference weight of the data-spreading
    scheme falls inside a group of 20 weights that are equal in exact arithmetic.  The literal >= picks a subset that
    depends on the last bits of the FBMC matrix (the oracle's FFT-built D0 and the product's closed-form D0 differ by
    1e-14 and already disagree); the exact-arithmetic reading (whole group, `TieTolerance`) makes oracle and product
    build the SAME precoder, bit for bit in its sparsity pattern."""
    from oracle.iic import ImaginaryInterferenceCancellationAtPilotPosition as RefIIC, TIE_RTOL
    S = ds_default
    pm_o, pm_f, pm_aux = DoublySelectiveSimulation._pilot_matrices(24, 1)
    D0_prod = FBMC(24, 30, 15e3, 15e3 * 24, 0, False, "Hermite-OQAM", 8, 0, True).GetFBMCMatrix()
    assert 0 < np.max(np.abs(D0_prod - S["D0"])) < 1e-12
    # the tie group: values 16..31 (1-based) of the sorted pattern are one weight (0.036858); both thresholds, the
    # 21st (Coding) and the 29th (Auxiliary) value, are inside it
    from oracle.iic import _interference_matrix
    srt = np.sort(np.abs(_interference_matrix(S["D0"], 24, 30)).reshape(-1))[::-1]
    assert np.ptp(srt[15:31]) < 1e-12 and srt[31] < 0.3 * srt[30] and srt[14] > 5 * srt[15]
    assert IIC.TieTolerance == TIE_RTOL == 1e-9
    for method, pm, n, off in (("Auxiliary", pm_aux, 28, 4.685), ("Coding", pm_f, 20, 4)):
        a = RefIIC(method, pm, S["D0"], n, off)
        b = IIC(method, pm, D0_prod, n, off)
        assert np.array_equal(a.PrecodingMatrix != 0, b.PrecodingMatrix != 0)
        assert np.max(np.abs(a.PrecodingMatrix - b.PrecodingMatrix)) < 1e-13
        tags = np.asarray(b.ConsideredInterferenceMatrix).reshape(-1, order="F")
        assert all(np.sum(tags == -p) == 16 for p in range(1, 17))           # the whole group: 16 interferers per pilot
    # literal comparison: rounding-noise dependent (documented)
    lit_ref = RefIIC("Coding", pm_f, S["D0"], 20, 4, tie_rtol=0.0)
    lit_prod = IIC("Coding", pm_f, D0_prod, 20, 4, TieTolerance=0.0)
    n_ref = np.count_nonzero(lit_ref.PrecodingMatrix)
    n_prod = np.count_nonzero(lit_prod.PrecodingMatrix)
    assert n_ref != n_prod or not np.array_equal(lit_ref.PrecodingMatrix != 0, lit_prod.PrecodingMatrix != 0)
    # paper geometry (DS.m:42-46): the threshold is the smallest member of its group, both readings agree
    fb = FBMC(24, 60, 15e3, 15e3 * 14 * 14, 0, False, "Hermite-OQAM", 8, 0, True)
    D0p = fb.GetFBMCMatrix()
    _, pm_f2, pm_aux2 = DoublySelectiveSimulation._pilot_matrices(24, 2)
    for method, pm, n, off in (("Auxiliary", pm_aux2, 28, 4.685), ("Coding", pm_f2, 20, 4)):
        x = IIC(method, pm, D0p, n, off, TieTolerance=0.0)
        y = IIC(method, pm, D0p, n, off)
        assert np.array_equal(x.PrecodingMatrix, y.PrecodingMatrix)
