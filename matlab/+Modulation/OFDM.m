classdef OFDM < handle
    % Drop-in for the reference's Modulation.OFDM (+Modulation/OFDM.m): same eight constructor arguments, the
    % properties the scripts read (Nr.Subcarriers / MCSymbols / SamplesTotal, PHY.TimeSpacing / SamplingRate /
    % SubcarrierSpacing / dt, Implementation.*), and Modulation / Demodulation / GetTXMatrix / GetRXMatrix
    % (DoublySelectiveChannelEstimation.m:67-82,193-195; SimpleVersion_DoublyFlat.m:118-135).
    % Modulation / Demodulation run on a B200 in their FFT form (chest_modulate_fft / chest_demodulate_fft: zero-padded
    % IFFT + cyclic prefix + zero guard, OFDM.m:153-181); the matrices are written in closed form on the host.
    % NOT EXECUTED in this repository's CI (no MATLAB / Octave); mirror of chest_b200.Modulation.OFDM.
    properties (SetAccess = private)
        Nr
        PHY
        Implementation
    end
    properties (Access = private)
        Handle = []
        G = []
    end
    methods
        function obj = OFDM(varargin)
            if numel(varargin) == 0
                varargin = {24, 14, 15e3, 15e3 * 24 * 14, 0, false, 1 / (14 * 15e3), 0};
            elseif numel(varargin) ~= 8
                error('Number of input variables must be either 0 (default values) or 8');
            end
            [obj.Nr.Subcarriers, obj.Nr.MCSymbols, obj.PHY.SubcarrierSpacing, obj.PHY.SamplingRate, ...
                obj.PHY.IntermediateFrequency, obj.PHY.TransmitRealSignal, obj.PHY.CyclicPrefixLength, ...
                obj.PHY.ZeroGuardTimeLength] = deal(varargin{:});
            if obj.PHY.TransmitRealSignal
                error('PHY.TransmitRealSignal == true is not supported by the B200 build');
            end
            obj.SetDependentParameters;
        end

        function SetDependentParameters(obj)
            fs = obj.PHY.SamplingRate;
            if mod(round(fs / obj.PHY.SubcarrierSpacing * 1e5) / 1e5, 1) ~= 0
                obj.PHY.SubcarrierSpacing = fs / round(fs / obj.PHY.SubcarrierSpacing);
                disp('Sampling rate must be a multiple of the subcarrier spacing!');
            end
            F = obj.PHY.SubcarrierSpacing;
            if mod(round(obj.PHY.IntermediateFrequency / F * 1e5) / 1e5, 1) ~= 0
                obj.PHY.IntermediateFrequency = round(obj.PHY.IntermediateFrequency / F) * F;
                disp('The intermediate frequency must be a multiple of the subcarrier spacing!');
            end
            if fs < obj.Nr.Subcarriers * F
                error('Sampling theorem is not fullfilled: sampling rate must be higher than the number of subcarriers times subcarrier spacing');
            end
            if abs(mod(round(obj.PHY.CyclicPrefixLength * fs * 1e5) / 1e5, 1)) ~= 0
                obj.PHY.CyclicPrefixLength = round(obj.PHY.CyclicPrefixLength * fs) / fs;
                disp('The length of the cyclic prefix times the sampling rate must be an integer!');
            end
            obj.Implementation.CyclicPrefix = round(obj.PHY.CyclicPrefixLength * fs);
            obj.Implementation.ZeroGuardSamples = round(obj.PHY.ZeroGuardTimeLength * fs);
            obj.Implementation.FFTSize = round(fs / F);
            obj.Implementation.TimeSpacing = obj.Implementation.FFTSize + obj.Implementation.CyclicPrefix;
            obj.Implementation.IntermediateFrequency = round(obj.PHY.IntermediateFrequency / F);
            obj.Implementation.NormalizationFactor = sqrt(fs^2 / F^2 / obj.Nr.Subcarriers);
            obj.PHY.dt = 1 / fs;
            obj.PHY.TimeSpacing = obj.Implementation.TimeSpacing * obj.PHY.dt;
            obj.Nr.SamplesTotal = obj.Nr.MCSymbols * obj.Implementation.TimeSpacing + 2 * obj.Implementation.ZeroGuardSamples;
            obj.ReleaseDevice;
            obj.G = [];
        end

        function TransmitSignal = Modulation(obj, DataSymbols)
            % L x K (x pages) data symbols -> N x 1 (x pages) samples, on the device
            L = obj.Nr.Subcarriers;  K = obj.Nr.MCSymbols;
            x = reshape(DataSymbols, L * K, []);
            TransmitSignal = chest_mex('modulate_fft', obj.Device, 1, x, obj.Nr.SamplesTotal);
        end

        function ReceivedSymbols = Demodulation(obj, ReceivedSignal)
            L = obj.Nr.Subcarriers;  K = obj.Nr.MCSymbols;
            y = chest_mex('demodulate_fft', obj.Device, 1, ReceivedSignal, L * K);
            ReceivedSymbols = reshape(y, L, K, []);
        end

        function TXMatrix = GetTXMatrix(obj)
            % G (N x L K): subcarrier l of symbol k is a complex exponential over FFTSize samples preceded by its
            % cyclic prefix -- the closed form of calling Modulation once per unit vector (OFDM.m:184-203)
            if isempty(obj.G)
                L = obj.Nr.Subcarriers;  K = obj.Nr.MCSymbols;  N = obj.Nr.SamplesTotal;  I = obj.Implementation;
                t = (0:I.TimeSpacing - 1).' - I.CyclicPrefix;
                Base = (I.NormalizationFactor / I.FFTSize) * exp(2j * pi * mod(t, I.FFTSize) * (I.IntermediateFrequency + (0:L - 1)) / I.FFTSize);
                obj.G = zeros(N, L * K);
                for k = 0:K - 1
                    obj.G(I.ZeroGuardSamples + k * I.TimeSpacing + (1:I.TimeSpacing), k * L + (1:L)) = Base;
                end
            end
            TXMatrix = obj.G;
        end

        function RXMatrix = GetRXMatrix(obj)
            % scaled G' with the cyclic-prefix samples ignored (OFDM.m:205-218)
            I = obj.Implementation;
            RXMatrix = obj.GetTXMatrix' * (obj.Nr.Subcarriers * obj.PHY.SubcarrierSpacing / obj.PHY.SamplingRate);
            idx = bsxfun(@plus, I.ZeroGuardSamples + (1:I.CyclicPrefix).', (0:obj.Nr.MCSymbols - 1) * I.TimeSpacing);
            RXMatrix(:, idx(:)) = 0;
        end

        function delete(obj)
            obj.ReleaseDevice;
        end
    end
    methods (Access = private)
        function h = Device(obj)
            if isempty(obj.Handle)
                I = obj.Implementation;
                obj.Handle = chest_mex('create', 0);
                chest_mex('set_modem', obj.Handle, 1, 1, obj.Nr.Subcarriers, obj.Nr.MCSymbols, I.FFTSize, ...
                    I.IntermediateFrequency + (0:obj.Nr.Subcarriers - 1), I.TimeSpacing, 1, I.CyclicPrefix, I.ZeroGuardSamples, ...
                    [], [], I.NormalizationFactor, obj.PHY.SubcarrierSpacing);
            end
            h = obj.Handle;
        end
        function ReleaseDevice(obj)
            if ~isempty(obj.Handle), chest_mex('destroy', obj.Handle); obj.Handle = []; end
        end
    end
end
