"""B200-native Monte-Carlo hot path of rnissel/Channel-Estimation (host side).

The package directory name contains a hyphen; import it with
``importlib.import_module("channel-estimation_b200")`` or through the root-level alias module
``chest_b200``.  Everything numerical on the hot path runs in ``libchest_b200.so`` (CUDA, sm_100a);
there is no CPU fallback.

Sub-packages mirror the reference's MATLAB packages:
  Channel.FastFading, Modulation.{FBMC,OFDM,SignalConstellation},
  ChannelEstimation.{PilotSymbolAidedChannelEstimation,ImaginaryInterferenceCancellationAtPilotPosition}
"""
from . import _lib, context, modulation, channel, estimation, simulation      # noqa: F401
from .context import DeviceContext, ChestError                                 # noqa: F401
from .simulation import DoublySelectiveSimulation                              # noqa: F401


class Channel:                                   # +Channel
    FastFading = channel.FastFading


class Modulation:                                # +Modulation
    FBMC = modulation.FBMC
    OFDM = modulation.OFDM
    SignalConstellation = modulation.SignalConstellation


class ChannelEstimation:                         # +ChannelEstimation
    PilotSymbolAidedChannelEstimation = estimation.PilotSymbolAidedChannelEstimation
    ImaginaryInterferenceCancellationAtPilotPosition = estimation.ImaginaryInterferenceCancellationAtPilotPosition
