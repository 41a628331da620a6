classdef FastFading < handle
    % Drop-in for the reference's Channel.FastFading (+Channel/FastFading.m): same nine constructor arguments,
    % same properties (PHY, Nr, Implementation, ImpulseResponse) and the methods the reference scripts call
    % (DoublySelectiveChannelEstimation.m:176-187, 352, 381; SimpleVersion_DoublyFlat.m does not use it).
    % NewRealization / GetConvolutionMatrix / Convolution run on a B200 through chest_mex (C ABI of
    % include/chest_b200.h); the constructor tables and the correlation functions are host-side setup.
    %
    %   ChannelModel = Channel.FastFading(SamplingRate, 'VehicularA', N, fD, 'Jakes', 200, 1, 1, true);
    %   ChannelModel.NewRealization;  H = ChannelModel.GetConvolutionMatrix{1};  r = ChannelModel.Convolution(s);
    %
    % NOT EXECUTED in this repository's CI (no MATLAB / Octave in the image); the Python mirror
    % chest_b200.Channel.FastFading implements the same logic and is what the tests run.  See INTEGRATION.md.
    properties (SetAccess = private)
        PHY
        Nr
        Implementation
        ImpulseResponse
    end
    properties (Access = private)
        Handle = []
        Seed = 0
        Count = 0
    end
    methods
        function obj = FastFading(SamplingRate, PowerDelayProfile, SamplesTotal, MaximumDopplerShift, ...
                DopplerModel, Paths, nTxAntennas, nRxAntennas, WarningIfSampleRateDoesNotMatch)
            if nargin < 7, nTxAntennas = 1; end
            if nargin < 8, nRxAntennas = 1; end
            if nargin < 9, WarningIfSampleRateDoesNotMatch = false; end
            obj.PHY.SamplingRate = SamplingRate;
            obj.PHY.dt = 1 / SamplingRate;
            obj.PHY.MaximumDopplerShift = MaximumDopplerShift;
            obj.PHY.DopplerModel = DopplerModel;
            obj.Nr.SamplesTotal = SamplesTotal;
            obj.Nr.txAntennas = nTxAntennas;
            obj.Nr.rxAntennas = nRxAntennas;
            obj.Nr.Paths = Paths;
            obj.Implementation.PowerDelayProfile = PowerDelayProfile;
            obj.Implementation.UseDiscreteDopplerSpectrum = false;
            dt = obj.PHY.dt;
            if ischar(PowerDelayProfile)
                % relative power (dB) over delay (s): ITU / 3GPP tables (the reference keeps the same tables, FF.m:56-107)
                [PowerdB, Delay] = Channel.FastFading.DelayProfileTable(PowerDelayProfile);
                IndexDelays = round(Delay / dt) + 1;
                if WarningIfSampleRateDoesNotMatch && (sum(abs(rem(Delay, dt))) > 0)
                    disp('Sampling rate does not match the predefined delays of the channel model!');
                end
                pdp = accumarray(IndexDelays(:), 10.^(PowerdB(:) / 10)).';      % taps that share a sample add up
                obj.PHY.DesiredPowerDelayProfiledB = [PowerdB; Delay];
            else
                pdp = PowerDelayProfile(:).';
            end
            obj.PHY.PowerDelayProfile = pdp;
            obj.Implementation.PowerDelayProfileNormalized = pdp(:) / sum(pdp);
            obj.Implementation.IndexDelayTaps = find(pdp(:));
            Discrete = strncmp(DopplerModel, 'Discrete', 8);
            if Discrete && MaximumDopplerShift > 0 && MaximumDopplerShift / (SamplingRate / SamplesTotal) <= 0.5
                disp('Discrete Doppler spectrum: The velocity is so low, that it is set to zero.');
                obj.PHY.MaximumDopplerShift = 0;
            end
            if Discrete && obj.PHY.MaximumDopplerShift > 0
                obj.Implementation.UseDiscreteDopplerSpectrum = true;
            end
            if ~any(strcmp(DopplerModel, {'Jakes', 'Uniform', 'Discrete-Jakes', 'Discrete-Uniform'})) && MaximumDopplerShift > 0
                error('Doppler spectrum not supported');
            end
            % ---- device context: one realization slot per antenna link (index rx + nRx * (tx - 1))
            obj.Handle = chest_mex('create', 0);
            if obj.PHY.MaximumDopplerShift > 0
                model = find(strcmp(DopplerModel, {'Jakes', 'Uniform', 'Discrete-Jakes', 'Discrete-Uniform'})) - 1;
            else
                model = 0;       % time-invariant: realizations are uploaded (set_impulse_response)
            end
            chest_mex('set_channel', obj.Handle, SamplesTotal, obj.Implementation.PowerDelayProfileNormalized, ...
                obj.PHY.MaximumDopplerShift, dt, Paths, model);
            chest_mex('finalize', obj.Handle, nTxAntennas * nRxAntennas);
            obj.NewRealization;
        end

        function SetSeed(obj, Seed)
            % realizations are drawn by the library's counter-based generator keyed by (seed, call count)
            obj.Seed = Seed;  obj.Count = 0;
        end

        function NewRealization(obj)
            % FF.m:194-250
            N = obj.Nr.SamplesTotal;  nL = obj.Nr.txAntennas * obj.Nr.rxAntennas;
            pdp = obj.Implementation.PowerDelayProfileNormalized;  Lt = numel(pdp);
            if ischar(obj.Implementation.PowerDelayProfile) && strcmp(obj.Implementation.PowerDelayProfile, 'AWGN')
                obj.ImpulseResponse = ones(1, 1, obj.Nr.rxAntennas, obj.Nr.txAntennas);
                chest_mex('set_impulse_response', obj.Handle, nL, repmat([1, zeros(1, Lt - 1)], [N 1 nL]));
            elseif obj.PHY.MaximumDopplerShift > 0
                chest_mex('new_realization', obj.Handle, nL, obj.Seed, obj.Count * nL);
                h = zeros(N, Lt, obj.Nr.rxAntennas, obj.Nr.txAntennas);
                for iTx = 1:obj.Nr.txAntennas
                    for iRx = 1:obj.Nr.rxAntennas
                        h(:, :, iRx, iTx) = chest_mex('impulse_response', obj.Handle, (iRx - 1) + obj.Nr.rxAntennas * (iTx - 1), N, Lt);
                    end
                end
                obj.ImpulseResponse = h;
            else
                % block fading: one complex normal per tap and link, constant over the block (FF.m:241-248)
                h = bsxfun(@times, sqrt(pdp.' / 2), randn(1, Lt, obj.Nr.rxAntennas, obj.Nr.txAntennas) + 1j * randn(1, Lt, obj.Nr.rxAntennas, obj.Nr.txAntennas));
                obj.ImpulseResponse = h;
                hl = reshape(h, 1, Lt, nL);
                chest_mex('set_impulse_response', obj.Handle, nL, repmat(hl, [N 1 1]));
            end
            obj.Count = obj.Count + 1;
        end

        function convolvedSignal = Convolution(obj, signal)
            % FF.m:253-274: r(:,iRx) = sum_iTx H{iRx,iTx} * s(:,iTx); the banded H is applied on the device
            N = size(signal, 1);
            convolvedSignal = zeros(N, obj.Nr.rxAntennas);
            for iTx = 1:obj.Nr.txAntennas
                for iRx = 1:obj.Nr.rxAntennas
                    convolvedSignal(:, iRx) = convolvedSignal(:, iRx) + ...
                        chest_mex('convolve', obj.Handle, (iRx - 1) + obj.Nr.rxAntennas * (iTx - 1), signal(:, iTx), N);
                end
            end
        end

        function ConvolutionMatrix = GetConvolutionMatrix(obj)
            % FF.m:276-295: nRx x nTx cell of sparse N x N matrices, H(r, r-m) = h(r, m+1)
            ConvolutionMatrix = cell(obj.Nr.rxAntennas, obj.Nr.txAntennas);
            for iTx = 1:obj.Nr.txAntennas
                for iRx = 1:obj.Nr.rxAntennas
                    c = chest_mex('convolution_matrix', obj.Handle, (iRx - 1) + obj.Nr.rxAntennas * (iTx - 1), obj.Nr.SamplesTotal);
                    ConvolutionMatrix{iRx, iTx} = c{1};
                end
            end
        end

        function ChannelTransferFunction = GetTransferFunction(obj, TimePos, FFTSize, ActiveSubcarrier)
            % FF.m:296-318
            if obj.PHY.MaximumDopplerShift == 0, TimePos = ones(numel(TimePos), 1); end
            ChannelTransferFunction = zeros(FFTSize, numel(TimePos), obj.Nr.rxAntennas, obj.Nr.txAntennas);
            for iTx = 1:obj.Nr.txAntennas
                for iRx = 1:obj.Nr.rxAntennas
                    taps = obj.ImpulseResponse(TimePos, :, iRx, iTx).';
                    ChannelTransferFunction(:, :, iRx, iTx) = fft([taps; zeros(FFTSize - size(taps, 1), numel(TimePos))], [], 1);
                end
            end
            if nargin > 3, ChannelTransferFunction = ChannelTransferFunction(ActiveSubcarrier, :, :, :); end
        end

        function [TimeCorrelation, Time] = GetTimeCorrelation(obj)
            % FF.m:321-340: 2N-1 lags, lag 0 at index N
            N = obj.Nr.SamplesTotal;
            Time = ((1:2 * N - 1) - N) * obj.PHY.dt;
            if obj.PHY.MaximumDopplerShift > 0
                if any(strcmp(obj.PHY.DopplerModel, {'Jakes', 'Discrete-Jakes'}))
                    TimeCorrelation = besselj(0, 2 * pi * obj.PHY.MaximumDopplerShift * Time);
                else
                    x = 2 * obj.PHY.MaximumDopplerShift * Time;
                    TimeCorrelation = ones(size(x));
                    nz = x ~= 0;
                    TimeCorrelation(nz) = sin(pi * x(nz)) ./ (pi * x(nz));
                end
            else
                TimeCorrelation = ones(1, 2 * N - 1);
            end
        end

        function [FrequencyCorrelation, Frequency] = GetFrequencyCorrelation(obj)
            % FF.m:342-351
            N = obj.Nr.SamplesTotal;  p = obj.Implementation.PowerDelayProfileNormalized;
            FrequencyCorrelation = circshift(fft([p; zeros(N - numel(p), 1)]), [ceil(N / 2) 1]);
            Frequency = ((1:N) - ceil(N / 2) - 1) / (N * obj.PHY.dt);
        end

        function MeanDelay = GetMeanDelay(obj)
            p = obj.Implementation.PowerDelayProfileNormalized.';
            MeanDelay = sum((0:numel(p) - 1) * obj.PHY.dt .* p);
        end

        function RmsDelaySpread = GetRmsDelaySpread(obj)
            p = obj.Implementation.PowerDelayProfileNormalized.';
            Tau = (0:numel(p) - 1) * obj.PHY.dt;
            RmsDelaySpread = sqrt(sum(Tau.^2 .* p) - obj.GetMeanDelay^2);
        end

        function CorrelationMatrix = GetCorrelationMatrix(obj)
            % R_vecH = E{H(:) H(:)'} of the vectorised convolution matrix (FF.m:366-407), N^2 x N^2 sparse:
            % entries of the same delay tap m are correlated with the time correlation of their row distance and
            % weighted by the tap power; different taps are uncorrelated (WSSUS).  The linear index of tap m in row
            % block r is (r-1)(N+1) + m (it runs into the next column for r + m - 1 > N, exactly as the reference's
            % index map does; indices beyond N^2 are dropped).
            % The Tier-2 path never needs this matrix: chest_setup_correlations forms the pilot correlations from
            % GetTimeCorrelation and the power delay profile directly on the device (ChestB200.Simulation).
            N = obj.Nr.SamplesTotal;
            p = obj.Implementation.PowerDelayProfileNormalized;
            Rt = obj.GetTimeCorrelation;
            [c, r] = meshgrid(1:N, 1:N);
            Toep = Rt(N + r - c);
            CorrelationMatrix = sparse(N^2, N^2);
            for m = find(p(:)).'
                idx = (0:N - 1).' * (N + 1) + m;
                keep = idx <= N^2;
                [J, I] = meshgrid(idx(keep), idx(keep));
                T = Toep(keep, keep) * p(m);
                CorrelationMatrix = CorrelationMatrix + sparse(I(:), J(:), T(:), N^2, N^2);
            end
        end

        function delete(obj)
            if ~isempty(obj.Handle), chest_mex('destroy', obj.Handle); obj.Handle = []; end
        end
    end

    methods (Static)
        function [PowerdB, Delay] = DelayProfileTable(Name)
            % named power delay profiles: relative power in dB, delay in seconds
            if strncmp(Name, 'TDL', 3)
                % 3GPP TR 38.900 tapped delay lines 'TDL-A_<rms delay spread>ns' (normalised delays scaled by the spread)
                us = strfind(Name, '_');  ns = strfind(Name, 'ns');
                Spread = str2double(Name(us + 1:ns - 1)) * 1e-9;
                switch Name(1:5)
                    case 'TDL-A'
                        PowerdB = [-13.4 0 -2.2 -4 -6 -8.2 -9.9 -10.5 -7.5 -15.9 -6.6 -16.7 -12.4 -15.2 -10.8 -11.3 -12.7 -16.2 -18.3 -18.9 -16.6 -19.9 -29.7];
                        Rel = [0.0000 0.3819 0.4025 0.5868 0.4610 0.5375 0.6708 0.5750 0.7618 1.5375 1.8978 2.2242 2.1718 2.4942 2.5119 3.0582 4.0810 4.4579 4.5695 4.7966 5.0066 5.3043 9.6586];
                    case 'TDL-B'
                        PowerdB = [0 -2.2 -4 -3.2 -9.8 -1.2 -3.4 -5.2 -7.6 -3 -8.9 -9 -4.8 -5.7 -7.5 -1.9 -7.6 -12.2 -9.8 -11.4 -14.9 -9.2 -11.3];
                        Rel = [0.0000 0.1072 0.2155 0.2095 0.2870 0.2986 0.3752 0.5055 0.3681 0.3697 0.5700 0.5283 1.1021 1.2756 1.5474 1.7842 2.0169 2.8294 3.0219 3.6187 4.1067 4.2790 4.7834];
                    case 'TDL-C'
                        PowerdB = [-4.4 -1.2 -3.5 -5.2 -2.5 0 -2.2 -3.9 -7.4 -7.1 -10.7 -11.1 -5.1 -6.8 -8.7 -13.2 -13.9 -13.9 -15.8 -17.1 -16 -15.7 -21.6 -22.8];
                        Rel = [0 0.2099 0.2219 0.2329 0.2176 0.6366 0.6448 0.6560 0.6584 0.7935 0.8213 0.9336 1.2285 1.3083 2.1704 2.7105 4.2589 4.6003 5.4902 5.6077 6.3065 6.6374 7.0427 8.6523];
                    otherwise
                        error('Power delay profile model not supported!');
                end
                Delay = Spread * Rel;
                return;
            end
            switch Name
                case {'Flat', 'AWGN'}
                    PowerdB = 0;  Delay = 0;
                case 'PedestrianA'
                    PowerdB = [0 -9.7 -19.2 -22.8];  Delay = [0 110 190 410] * 1e-9;
                case 'PedestrianB'
                    PowerdB = [0 -0.9 -4.9 -8 -7.8 -23.9];  Delay = [0 200 800 1200 2300 3700] * 1e-9;
                case 'VehicularA'
                    PowerdB = [0 -1 -9 -10 -15 -20];  Delay = [0 310 710 1090 1730 2510] * 1e-9;
                case 'VehicularB'
                    PowerdB = [-2.5 0 -12.8 -10 -25.2 -16];  Delay = [0 300 8900 12900 17100 20000] * 1e-9;
                case 'ExtendedPedestrianA'
                    PowerdB = [0 -1 -2 -3 -8 -17.2 -20.8];  Delay = [0 30 70 90 110 190 410] * 1e-9;
                case 'ExtendedVehicularA'
                    PowerdB = [0 -1.5 -1.4 -3.6 -0.6 -9.1 -7 -12 -16.9];  Delay = [0 30 150 310 370 710 1090 1730 2510] * 1e-9;
                otherwise
                    error('Power delay profile model not supported!');
            end
        end
    end
end
