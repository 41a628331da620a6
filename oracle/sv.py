"""Oracle restatement of the loop body of SimpleVersion_DoublyFlat.m (SV.m:89-176): one (repetition, SNR point) pass.
Test infrastructure only -- see oracle/__init__.py.

The interpolation matrices (ChannelEstimation_*.GetInterpolationMatrix, PSACE.m:171-184) are inputs: the reference builds
them with MATLAB's closed-source scatteredInterpolant, whose triangulation / extrapolation choices are implementation
defined; everything else follows the script line by line with the oracle's FFT modem."""
import numpy as np

from .fbmc import FBMC
from .ofdm import OFDM
from .iic import ImaginaryInterferenceCancellationAtPilotPosition as IIC
from .signal_constellation import SignalConstellation


def sv_setup(pilot_matrix_fbmc, aux_matrix_fbmc, pilot_matrix_ofdm, interp_fbmc, interp_ofdm):
    """SV.m:12-82 with the pilot matrices / interpolation matrices handed in."""
    S = {}
    S["fbmc"] = FBMC(12, 30, 15e3, 15e3 * 14 * 12, 15e3 * 20, False, "Hermite-OQAM", 8, 0, True)          # SV.m:17-28
    S["ofdm"] = OFDM(12, 15, 15e3, 15e3 * 14 * 12, 15e3 * 20, False, 0, (8 - 1 / 2) * 1 / 15e3 * 1 / 2)   # SV.m:31-40
    S["PAM"], S["QAM"] = SignalConstellation(4, "PAM"), SignalConstellation(16, "QAM")                     # SV.m:43-44
    D0 = S["fbmc"].GetFBMCMatrix()
    S["aux"] = IIC("Auxiliary", aux_matrix_fbmc, D0, 16, 2)                                                # SV.m:71-75
    S["cod"] = IIC("Coding", pilot_matrix_fbmc, D0, 16, 2)                                                 # SV.m:76-80
    S["pm_f"], S["pm_aux"], S["pm_o"] = (np.asarray(m).reshape(-1, order="F") for m in
                                         (pilot_matrix_fbmc, aux_matrix_fbmc, pilot_matrix_ofdm))
    S["interp_f"], S["interp_o"] = np.asarray(interp_fbmc), np.asarray(interp_ofdm)
    S["P_f"], S["P_o"] = int(np.sum(S["pm_f"] == 1)), int(np.sum(S["pm_o"] == 1))
    S["N"] = S["fbmc"].Nr["SamplesTotal"]
    return S


def sv_pn(S, snr_db):
    o = S["ofdm"]
    return o.PHY["SamplingRate"] / (o.PHY["SubcarrierSpacing"] * o.Nr["Subcarriers"]) * 10 ** (-snr_db / 10)   # SV.m:92


def sv_new_draws(S, rng):
    """Draws of one body in the order SV.m:95-126 consumes them."""
    nb_a, nb_c = S["aux"].NrDataSymbols * 2, S["cod"].NrDataSymbols * 2
    nb_o = (12 * 15 - S["P_o"]) * 4
    N = S["N"]
    return dict(bits_aux=rng.integers(0, 2, nb_a).astype(np.uint8), bits_cod=rng.integers(0, 2, nb_c).astype(np.uint8),
                bits_ofdm=rng.integers(0, 2, nb_o).astype(np.uint8),
                pil_idx_fbmc=rng.integers(0, 4, S["P_f"]).astype(np.int32), pil_idx_ofdm=rng.integers(0, 16, S["P_o"]).astype(np.int32),
                h=np.sqrt(0.5) * (rng.standard_normal() + 1j * rng.standard_normal()),
                noise_fbmc=rng.standard_normal(N) + 1j * rng.standard_normal(N),
                noise_ofdm=rng.standard_normal(N) + 1j * rng.standard_normal(N))


def sv_body(S, d, pn):
    """SV.m:95-169 for one body.  Returns the five bit-error counts (aux, cod, FBMC perfect, OFDM, OFDM perfect)."""
    fb, of, pam, qam, aux, cod = S["fbmc"], S["ofdm"], S["PAM"], S["QAM"], S["aux"], S["cod"]
    pmf, pma, pmo = S["pm_f"], S["pm_aux"], S["pm_o"]
    P = S["P_f"]
    xP = pam.SymbolMapping[d["pil_idx_fbmc"]]; xP = xP / np.abs(xP)                                       # SV.m:105-106
    xPo = qam.SymbolMapping[d["pil_idx_ofdm"]]; xPo = xPo / np.abs(xPo)                                   # SV.m:107-108
    x_aux = (aux.PrecodingMatrix @ np.concatenate([xP, pam.Bit2Symbol(d["bits_aux"])])).reshape(12, 30, order="F")   # SV.m:111
    x_cod = (cod.PrecodingMatrix @ np.concatenate([xP, pam.Bit2Symbol(d["bits_cod"])])).reshape(12, 30, order="F")   # SV.m:112
    x_o = np.zeros(180, dtype=complex)
    x_o[pmo == 1] = xPo; x_o[pmo == 0] = qam.Bit2Symbol(d["bits_ofdm"])                                   # SV.m:113-115
    s_aux, s_cod, s_o = fb.Modulation(x_aux), fb.Modulation(x_cod), of.Modulation(x_o.reshape(12, 15, order="F"))   # SV.m:118-120
    h = d["h"]
    n_f, n_o = np.sqrt(pn / 2) * d["noise_fbmc"], np.sqrt(pn / 2) * d["noise_ofdm"]                       # SV.m:125-126
    y_aux = fb.Demodulation(h * s_aux + n_f).reshape(-1, order="F")                                        # SV.m:128-135
    y_cod = fb.Demodulation(h * s_cod + n_f).reshape(-1, order="F")
    y_o = of.Demodulation(h * s_o + n_o).reshape(-1, order="F")
    hP_aux = y_aux[pmf == 1] / xP / np.sqrt(aux.PilotToDataPowerOffset * aux.DataPowerReduction)          # SV.m:138
    hP_cod = y_cod[pmf == 1] / xP / np.sqrt(cod.PilotToDataPowerOffset)                                    # SV.m:139
    hP_o = y_o[pmo == 1] / xPo                                                                             # SV.m:140
    h_aux, h_cod, h_o = S["interp_f"] @ hP_aux, S["interp_f"] @ hP_cod, S["interp_o"] @ hP_o              # SV.m:143-145
    eq_aux = np.real(y_aux[pma == 0] / h_aux[pma == 0] / np.sqrt(aux.DataPowerReduction))                  # SV.m:148
    Cd = cod.PrecodingMatrix[:, P:]
    eq_cod = np.real(Cd.conj().T @ (y_cod / h_cod))                                                        # SV.m:149
    eq_perf = np.real(Cd.conj().T @ (y_cod / h))                                                           # SV.m:150
    eq_o, eq_o_perf = y_o[pmo == 0] / h_o[pmo == 0], y_o[pmo == 0] / h                                     # SV.m:152-153
    return np.array([np.sum(pam.Symbol2Bit(eq_aux) != d["bits_aux"]), np.sum(pam.Symbol2Bit(eq_cod) != d["bits_cod"]),
                     np.sum(pam.Symbol2Bit(eq_perf) != d["bits_cod"]), np.sum(qam.Symbol2Bit(eq_o) != d["bits_ofdm"]),
                     np.sum(qam.Symbol2Bit(eq_o_perf) != d["bits_ofdm"])], dtype=np.int64)                 # SV.m:156-169
