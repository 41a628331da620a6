classdef SignalConstellation < handle
    % Drop-in for the reference's Modulation.SignalConstellation (+Modulation/SignalConstellation.m:24-101):
    % Gray-labelled M-PAM / M-QAM with unit mean power, SymbolMapping sorted by the bit word (first bit column = LSB),
    % BitMapping M x log2(M).  Bit2Symbol / Symbol2Bit / SymbolQuantization are the host versions (the hot path demaps
    % on the device: kernels.cuh demap_word; the tables are uploaded with chest_mex('set_constellation', ...)).
    % NOT EXECUTED in this repository's CI (no MATLAB / Octave); mirror of chest_b200.Modulation.SignalConstellation.
    properties (SetAccess = private)
        Method
        ModulationOrder
        BitMapping
        SymbolMapping
        Implementation
    end
    methods
        function obj = SignalConstellation(ModulationOrder, Method)
            obj.ModulationOrder = ModulationOrder;
            obj.Method = Method;
            M = ModulationOrder;
            switch Method
                case 'PAM'
                    Gray = Modulation.SignalConstellation.AxisBits(M);
                    Symbols = (2 * (1:M) - M - 1).';
                case 'QAM'
                    Ms = round(sqrt(M));
                    if Ms * Ms ~= M, error('QAM order must be a square'); end
                    Ax = Modulation.SignalConstellation.AxisBits(Ms);
                    nb = size(Ax, 2);
                    Level = 2 * (1:Ms) - Ms - 1;
                    % column-major grid: index = q + Ms * i; odd bit columns follow the I level, even ones the Q level
                    idx = (0:M - 1).';
                    ii = floor(idx / Ms) + 1;  qq = mod(idx, Ms) + 1;
                    Symbols = Level(ii).' + 1j * Level(qq).';
                    Gray = zeros(M, 2 * nb);
                    Gray(:, 2:2:end) = Ax(qq, :);
                    Gray(:, 1:2:end) = Ax(ii, :);
                otherwise
                    error('Signal constellation method must be QAM or PAM!');
            end
            Symbols = Symbols / sqrt(mean(abs(Symbols).^2));
            Words = Gray * (2.^(0:size(Gray, 2) - 1)).';
            [~, Order] = sort(Words);
            obj.SymbolMapping = Symbols(Order);
            obj.BitMapping = Gray(Order, :);
            obj.Implementation.BitsPerSymbol = size(Gray, 2);
        end

        function DataSymbols = Bit2Symbol(obj, BinaryStream)
            nb = obj.Implementation.BitsPerSymbol;
            Words = reshape(BinaryStream, nb, []).' * (2.^(0:nb - 1)).';
            DataSymbols = obj.SymbolMapping(Words + 1);
        end

        function EstimatedBitStream = Symbol2Bit(obj, EstimatedDataSymbols)
            Index = obj.Nearest(EstimatedDataSymbols);
            EstimatedBitStream = reshape(obj.BitMapping(Index, :).', [], 1);
        end

        function QuantizedDataSymbols = SymbolQuantization(obj, EstimatedDataSymbols)
            QuantizedDataSymbols = reshape(obj.SymbolMapping(obj.Nearest(EstimatedDataSymbols)), size(EstimatedDataSymbols));
        end
    end
    methods (Access = private)
        function Index = Nearest(obj, x)
            % nearest constellation point, first index on ties (min over the columns of the distance matrix)
            [~, Index] = min(abs(bsxfun(@minus, x(:), obj.SymbolMapping(:).')), [], 2);
        end
    end
    methods (Static)
        function t = AxisBits(n)
            % Gray labels of one amplitude axis: first bit splits the axis in halves, every further bit is the
            % previous one decimated by two and mirrored
            nb = round(log2(n));
            t = zeros(n, nb);
            t(1:n / 2, 1) = 1;
            for c = 2:nb
                half = t(1:2:end, c - 1);
                t(:, c) = [half; flipud(half)];
            end
        end
    end
end
