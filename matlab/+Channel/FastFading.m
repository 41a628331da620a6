classdef FastFading < handle
    % Drop-in for the reference's Channel.FastFading (same constructor arguments, same
    % properties, same methods used by DoublySelectiveChannelEstimation.m) whose
    % NewRealization / GetConvolutionMatrix / Convolution run on a B200 through chest_mex.
    % The power-delay-profile tables stay in the reference's own constructor code: pass the
    % reference object's Implementation.PowerDelayProfileNormalized (FastFading.m:129), or a
    % vector power delay profile, as PowerDelayProfile.
    % NOT EXECUTED in this repository's CI (no MATLAB / Octave in the image); see INTEGRATION.md.
    properties (SetAccess = private)
        PHY
        Nr
        Implementation
        ImpulseResponse
    end
    properties (Access = private)
        Handle
        Seed = 0
        Count = 0
    end
    methods
        function obj = FastFading(SamplingRate, PowerDelayProfile, SamplesTotal, MaximumDopplerShift, ...
                DopplerModel, Paths, nTxAntennas, nRxAntennas, ~)
            assert(nTxAntennas == 1 && nRxAntennas == 1, 'only 1x1 antennas are supported');
            assert(isnumeric(PowerDelayProfile), 'pass the normalised power delay profile vector');
            obj.PHY.SamplingRate = SamplingRate;  obj.PHY.dt = 1 / SamplingRate;
            obj.PHY.MaximumDopplerShift = MaximumDopplerShift;  obj.PHY.DopplerModel = DopplerModel;
            obj.Nr.SamplesTotal = SamplesTotal;  obj.Nr.Paths = Paths;
            obj.Nr.txAntennas = 1;  obj.Nr.rxAntennas = 1;
            obj.PHY.PowerDelayProfile = PowerDelayProfile(:).';
            obj.Implementation.PowerDelayProfileNormalized = PowerDelayProfile(:) / sum(PowerDelayProfile);
            obj.Implementation.IndexDelayTaps = find(PowerDelayProfile(:));
            model = find(strcmp(DopplerModel, {'Jakes', 'Uniform'})) - 1;
            obj.Handle = chest_mex('create', 0);
            chest_mex('set_channel', obj.Handle, SamplesTotal, obj.Implementation.PowerDelayProfileNormalized, ...
                MaximumDopplerShift, obj.PHY.dt, Paths, model);
            chest_mex('finalize', obj.Handle, 1);
            obj.NewRealization;
        end
        function NewRealization(obj)
            chest_mex('new_realization', obj.Handle, 1, obj.Seed, obj.Count);
            obj.Count = obj.Count + 1;
            obj.ImpulseResponse = chest_mex('impulse_response', obj.Handle, 0, obj.Nr.SamplesTotal, ...
                numel(obj.Implementation.PowerDelayProfileNormalized));
        end
        function ConvolutionMatrix = GetConvolutionMatrix(obj)
            ConvolutionMatrix = chest_mex('convolution_matrix', obj.Handle, 0, obj.Nr.SamplesTotal);
        end
        function convolvedSignal = Convolution(obj, signal)
            convolvedSignal = chest_mex('convolve', obj.Handle, 0, signal, size(signal, 1));
        end
        function delete(obj)
            if ~isempty(obj.Handle), chest_mex('destroy', obj.Handle); end
        end
    end
end
