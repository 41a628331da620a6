classdef FBMC < handle
    % Drop-in for the reference's Modulation.FBMC (+Modulation/FBMC.m), Hermite-OQAM polyphase path: same ten
    % constructor arguments, the properties the scripts read (Nr.*, PHY.*, PrototypeFilter.*, Implementation.*), and
    % Modulation / Demodulation / GetTXMatrix / GetRXMatrix / GetFBMCMatrix / GetInterferenceMatrix
    % (DoublySelectiveChannelEstimation.m:51-66,119,191-192; SimpleVersion_DoublyFlat.m:118-135).
    % Modulation / Demodulation run on a B200 in their FFT form (chest_modulate_fft / chest_demodulate_fft: polyphase
    % IFFT, prototype filter, overlap-add; FBMC.m:255-302); the matrices are written in closed form on the host.
    % Other prototype filters (PHYDYAS, RRC) and TransmitRealSignal are outside the accelerated path: keep the reference
    % class for those.
    % NOT EXECUTED in this repository's CI (no MATLAB / Octave); mirror of chest_b200.Modulation.FBMC.
    properties (SetAccess = private)
        Method
        Nr
        PHY
        PrototypeFilter
        Implementation
    end
    properties (Access = private)
        Handle = []
        G = []
    end
    methods
        function obj = FBMC(varargin)
            if numel(varargin) == 0
                varargin = {12, 30, 15e3, 12 * 15e3, 0, false, 'Hermite-OQAM', 8, 0, true};
            elseif numel(varargin) ~= 10
                error('Number of input variables must be either 0 (default values) or 10');
            end
            [obj.Nr.Subcarriers, obj.Nr.MCSymbols, obj.PHY.SubcarrierSpacing, obj.PHY.SamplingRate, ...
                obj.PHY.IntermediateFrequency, obj.PHY.TransmitRealSignal, obj.Method, ...
                obj.PrototypeFilter.OverlappingFactor, obj.Implementation.InitialPhaseShift, ...
                obj.Implementation.UsePolyphase] = deal(varargin{:});
            if ~strcmp(obj.Method, 'Hermite-OQAM')
                error('Method (prototype filter) "%s" is not supported by the B200 build', obj.Method);
            end
            if obj.PHY.TransmitRealSignal
                error('TransmitRealSignal = true is not supported by the B200 build');
            end
            obj.SetDependentParameters;
        end

        function SetDependentParameters(obj)
            fs = obj.PHY.SamplingRate;
            if mod(fs / (2 * obj.PHY.SubcarrierSpacing), 1) ~= 0
                obj.PHY.SubcarrierSpacing = fs / (2 * round(fs / (2 * obj.PHY.SubcarrierSpacing)));
                disp('Sampling Rate divided by (Subcarrier spacing times 2) must be must be an integer!');
            end
            F = obj.PHY.SubcarrierSpacing;
            if mod(obj.PHY.IntermediateFrequency / F, 1) ~= 0
                obj.PHY.IntermediateFrequency = round(obj.PHY.IntermediateFrequency / F) * F;
                disp('The intermediate frequency must be a multiple of the subcarrier spacing!');
            end
            if fs < obj.Nr.Subcarriers * F
                error('Sampling Rate must be higher: at least Number of Subcarriers times Subcarrier Spacing');
            end
            obj.PHY.dt = 1 / fs;
            obj.Implementation.TimeSpacing = round(fs / (2 * F));
            obj.PHY.TimeSpacing = obj.Implementation.TimeSpacing * obj.PHY.dt;
            obj.Implementation.FrequencySpacing = obj.PrototypeFilter.OverlappingFactor;
            obj.PrototypeFilter.TimeDomain = Modulation.FBMC.HermitePrototype(obj.PHY.TimeSpacing * 2, obj.PHY.dt, ...
                obj.PrototypeFilter.OverlappingFactor / 2);
            Np = numel(obj.PrototypeFilter.TimeDomain);
            obj.Nr.SamplesPrototypeFilter = Np;
            obj.Nr.SamplesTotal = Np + (obj.Nr.MCSymbols - 1) * obj.Implementation.TimeSpacing;
            obj.Implementation.FFTSize = round(Np / obj.Implementation.FrequencySpacing);
            obj.Implementation.IntermediateFrequency = round(obj.PHY.IntermediateFrequency / F);
            obj.Implementation.NormalizationFactor = sqrt(fs^2 / F^2 * obj.PHY.TimeSpacing / obj.Nr.Subcarriers);
            [k, l] = meshgrid(0:obj.Nr.MCSymbols - 1, 0:obj.Nr.Subcarriers - 1);
            obj.Implementation.PhaseShift = exp(1j * pi / 2 * (l + k)) * exp(1j * obj.Implementation.InitialPhaseShift);
            obj.ReleaseDevice;
            obj.G = [];
        end

        function TransmitSignal = Modulation(obj, DataSymbols)
            L = obj.Nr.Subcarriers;  K = obj.Nr.MCSymbols;
            x = reshape(DataSymbols, L * K, []);
            TransmitSignal = chest_mex('modulate_fft', obj.Device, 0, x, obj.Nr.SamplesTotal);
        end

        function ReceivedSymbols = Demodulation(obj, ReceivedSignal)
            L = obj.Nr.Subcarriers;  K = obj.Nr.MCSymbols;
            y = chest_mex('demodulate_fft', obj.Device, 0, ReceivedSignal, L * K);
            ReceivedSymbols = reshape(y, L, K, []);
        end

        function TXMatrix = GetTXMatrix(obj)
            % G (N x L K): column (l, k) is the prototype filter modulated to FFT bin b_l, delayed by k * TimeSpacing
            % and rotated by j^(l+k) -- the closed form of calling Modulation once per unit vector (FBMC.m:318-342)
            if isempty(obj.G)
                L = obj.Nr.Subcarriers;  K = obj.Nr.MCSymbols;  N = obj.Nr.SamplesTotal;  I = obj.Implementation;
                Np = obj.Nr.SamplesPrototypeFilter;
                Bins = sort(mod(I.IntermediateFrequency + (0:L - 1), I.FFTSize));
                t = (0:Np - 1).';
                Base = bsxfun(@times, obj.PrototypeFilter.TimeDomain(:) * (I.NormalizationFactor / I.FFTSize), ...
                    exp(2j * pi * t * Bins / I.FFTSize));
                Base = bsxfun(@times, Base, exp(1j * pi / 2 * (0:L - 1)) * exp(1j * I.InitialPhaseShift));
                obj.G = zeros(N, L * K);
                for k = 0:K - 1
                    obj.G(k * I.TimeSpacing + (1:Np), k * L + (1:L)) = Base * 1j^k;
                end
            end
            TXMatrix = obj.G;
        end

        function RXMatrix = GetRXMatrix(obj)
            RXMatrix = obj.GetTXMatrix' * (obj.Nr.Subcarriers / (obj.PHY.SamplingRate * obj.PHY.TimeSpacing));
        end

        function FBMCMatrix = GetFBMCMatrix(obj, FastCalculation)
            % D0 with y = D0 * x over a flat channel (FBMC.m:355-388).  The fast variant reads every entry off the
            % interference pattern of one impulse by (delta subcarrier, delta symbol), so entries with equal offsets are
            % bit-identical -- the pilot precoders' selection of "the N strongest interferers" relies on that.
            if nargin < 2, FastCalculation = true; end
            if ~FastCalculation
                FBMCMatrix = obj.GetRXMatrix * obj.GetTXMatrix;
                return;
            end
            L = obj.Nr.Subcarriers;  K = obj.Nr.MCSymbols;
            IM = obj.GetInterferenceMatrix;
            l = repmat((0:L - 1).', K, 1);
            k = kron((0:K - 1).', ones(L, 1));
            dl = bsxfun(@minus, l, l.');  dk = bsxfun(@minus, k, k.');
            D0 = IM(sub2ind(size(IM), dl + L, dk + K));
            TF = obj.PHY.TimeSpacing * obj.PHY.SubcarrierSpacing;
            FBMCMatrix = D0 .* exp(-1j * pi / 2 * (dl + dk)) .* exp(-1j * 2 * pi * TF * dk .* (bsxfun(@plus, l, dl / 2)));
        end

        function InterferenceMatrix = GetInterferenceMatrix(obj)
            % response of every (subcarrier, symbol) position to one impulse at (1, 1), mirrored to all four
            % quadrants of offsets (FBMC.m:390-400)
            L = obj.Nr.Subcarriers;  K = obj.Nr.MCSymbols;
            G_ = obj.GetTXMatrix;
            Y = reshape(obj.GetRXMatrix * G_(:, 1), L, K);
            [k, l] = meshgrid(0:K - 1, 0:L - 1);
            Y = Y .* exp(1j * pi / 2 * (l + k)) .* exp(-1j * pi * k .* (l / 2));
            InterferenceMatrix = [Y(end:-1:2, end:-1:2), Y(end:-1:2, :); Y(:, end:-1:2), Y];
        end

        function delete(obj)
            obj.ReleaseDevice;
        end
    end
    methods (Access = private)
        function h = Device(obj)
            if isempty(obj.Handle)
                I = obj.Implementation;  L = obj.Nr.Subcarriers;
                obj.Handle = chest_mex('create', 0);
                chest_mex('set_modem', obj.Handle, 0, 0, L, obj.Nr.MCSymbols, I.FFTSize, ...
                    sort(mod(I.IntermediateFrequency + (0:L - 1), I.FFTSize)), I.TimeSpacing, I.FrequencySpacing, 0, 0, ...
                    obj.PrototypeFilter.TimeDomain, I.PhaseShift, I.NormalizationFactor, obj.PHY.SubcarrierSpacing);
            end
            h = obj.Handle;
        end
        function ReleaseDevice(obj)
            if ~isempty(obj.Handle), chest_mex('destroy', obj.Handle); obj.Handle = []; end
        end
    end
    methods (Static)
        function p = HermitePrototype(T0, dt, OF)
            % Hermite prototype filter: weighted sum of the Hermite functions of order 0, 4, ..., 20 (Haas & Belfiore),
            % Horner evaluation of the even Hermite polynomials in z = (sqrt(2 pi) u)^2, unit energy (FBMC.m:629-647)
            n = round(2 * OF * T0 / dt);
            t = -(OF * T0) + (0:n - 1).' * dt;
            u = t / (T0 / sqrt(2));
            z = (sqrt(2 * pi) * u).^2;
            Coef = {1, [12 -48 16], [1680 -13440 13440 -3584 256], ...
                [665280 -7983360 13305600 -7096320 1520640 -135168 4096], ...
                [518918400 -8302694400 19372953600 -15498362880 5535129600 -984023040 89456640 -3932160 65536], ...
                [670442572800 -13408851456000 40226554368000 -42908324659200 21454162329600 -5721109954560 866834841600 -76205260800 3810263040 -99614720 1048576]};
            Weight = [1.412692577, -3.0145e-3, -8.8041e-6, -2.2611e-9, -4.4570e-15, 1.8633e-16];
            p = zeros(n, 1);
            for i = 1:numel(Coef)
                c = Coef{i};
                Poly = zeros(n, 1);
                for j = numel(c):-1:1
                    Poly = Poly .* z + c(j);
                end
                p = p + Weight(i) * Poly;
            end
            p = p .* exp(-pi * u.^2) / sqrt(T0);
            p = p / sqrt(sum(p.^2) * dt);
        end
    end
end
