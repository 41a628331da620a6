classdef ImaginaryInterferenceCancellationAtPilotPosition < handle
    % Drop-in for the reference's ChannelEstimation.ImaginaryInterferenceCancellationAtPilotPosition
    % (+ChannelEstimation/ImaginaryInterferenceCancellationAtPilotPosition.m:37-229): the FBMC precoding matrix that
    % cancels the imaginary interference at the pilot positions, by auxiliary symbols or by spreading the data.
    %   obj = ...(Method, PilotMatrix, FBMCMatrix, NrCanceledInterferersPerPilot, PilotToDataPowerOffset)
    % Same properties as the reference (PrecodingMatrix, NrDataSymbols, NrPilotSymbols, NrAuxiliarySymbols,
    % PilotToDataPowerOffset, DataPowerReduction, AuxiliaryToDataPowerOffset, SIR_dB, PilotMatrix,
    % ConsideredInterferenceMatrix, PostCodingChannelMatrix, NrTransmittedSymbols).  One-time host-side setup; its
    % PrecodingMatrix goes to the device through chest_mex('set_scheme', ...) (ChestB200.Simulation).
    % TieTolerance: the reference selects "the n strongest interferers" with a plain >= against the (n+1)-th largest
    % weight of the interference pattern; where that threshold falls inside a group of weights that are equal in exact
    % arithmetic, which members pass depends on their last bits.  TieTolerance > 0 (default 1e-9, relative) admits the
    % whole group -- the exact-arithmetic reading; 0 reproduces the literal comparison.  See DESIGN.md section 2.
    % NOT EXECUTED in this repository's CI (no MATLAB / Octave); mirror of chest_b200.ChannelEstimation.*.
    properties (SetAccess = private)
        Method
        PilotMatrix
        PrecodingMatrix
        NrDataSymbols
        NrPilotSymbols
        NrAuxiliarySymbols
        NrTransmittedSymbols
        PilotToDataPowerOffset
        DataPowerReduction
        AuxiliaryToDataPowerOffset
        SIR_dB
        ConsideredInterferenceMatrix
        PostCodingChannelMatrix
    end
    properties
        TieTolerance = 1e-9
    end
    methods
        function obj = ImaginaryInterferenceCancellationAtPilotPosition(Method, PilotMatrix, FBMCMatrix, ...
                NrCanceledInterferersPerPilot, PilotToDataPowerOffset, TieTolerance)
            if nargin > 5, obj.TieTolerance = TieTolerance; end
            [L, K] = size(PilotMatrix);
            LK = L * K;
            pm = PilotMatrix(:);
            pil = find(pm == 1);
            P = numel(pil);
            D0 = FBMCMatrix;
            obj.Method = Method;  obj.PilotMatrix = PilotMatrix;  obj.NrPilotSymbols = P;
            obj.PilotToDataPowerOffset = PilotToDataPowerOffset;  obj.NrTransmittedSymbols = LK;
            switch Method
                case 'Auxiliary'
                    dat = find(pm == 0);  aux = find(pm == -1);
                    nD = numel(dat);  nA = numel(aux);
                    Inv = pinv(D0(pil, aux));
                    C = zeros(LK, LK - nA);
                    C(aux, 1:P) = Inv * (eye(P) - D0(pil, pil));
                    C(aux, P + (1:nD)) = -Inv * D0(pil, dat);
                    C(sub2ind(size(C), pil, (1:P).')) = sqrt(PilotToDataPowerOffset);
                    C(sub2ind(size(C), dat, P + (1:nD).')) = 1;
                    if NrCanceledInterferersPerPilot > 0
                        [~, Tags] = obj.InterfererTags(D0, PilotMatrix, pil, NrCanceledInterferersPerPilot);
                        Keep = [Tags(pil); Tags(dat)] ~= 0;
                        C(aux, ~Keep) = 0;
                        obj.ConsideredInterferenceMatrix = reshape(Tags, L, K);
                    else
                        obj.ConsideredInterferenceMatrix = 'All';
                    end
                    obj.NrDataSymbols = nD;  obj.NrAuxiliarySymbols = nA;
                    obj.PostCodingChannelMatrix = nan;
                case 'Coding'
                    [Mask, Tags] = obj.InterfererTags(D0, PilotMatrix, pil, NrCanceledInterferersPerPilot);
                    if any(sum(Mask, 1) > 1)
                        error('Coding symbols must not overlap: The pilot-spacing is too small!');
                    end
                    Free = find(Tags == 0);
                    C = zeros(LK, LK - P);
                    C(sub2ind(size(C), pil, (1:P).')) = sqrt(PilotToDataPowerOffset);
                    C(sub2ind(size(C), Free, P + (1:numel(Free)).')) = 1;
                    Col0 = P + numel(Free);
                    for ip = 1:P
                        Pos = find(Tags == -ip);
                        Code = ChannelEstimation.ImaginaryInterferenceCancellationAtPilotPosition.SpreadingCode(D0(pil(ip), Pos));
                        C(Pos, Col0 + (1:size(Code, 2))) = Code;
                        Col0 = Col0 + size(Code, 2);
                    end
                    obj.NrDataSymbols = LK - 2 * P;  obj.NrAuxiliarySymbols = 0;
                    obj.ConsideredInterferenceMatrix = reshape(Tags, L, K);
                otherwise
                    error('Method must be  ''Auxiliary'' or ''Coding''!');
            end
            obj.DataPowerReduction = LK / sum(abs(C(:)).^2);
            C = C * sqrt(obj.DataPowerReduction);
            if strcmp(Method, 'Auxiliary')
                Power = sum(abs(C).^2, 2);
                obj.AuxiliaryToDataPowerOffset = mean(Power(pm == -1)) / mean(Power(pm == 0));
            else
                obj.AuxiliaryToDataPowerOffset = 0;
                obj.PostCodingChannelMatrix = abs(C').^2;
            end
            T = D0(pil, :) * C;
            Sig = abs(diag(T(:, 1:P))).^2;
            obj.SIR_dB = 10 * log10(Sig ./ (sum(abs(T).^2, 2) - Sig));
            obj.PrecodingMatrix = C;
        end
    end
    methods (Access = private)
        function [Mask, Tags] = InterfererTags(obj, D0, PilotMatrix, pil, NrCancel)
            % positions whose interference weight towards pilot p is among the NrCancel largest weights of the
            % interference pattern get tag -p, the pilots themselves +p
            [L, K] = size(PilotMatrix);  LK = L * K;
            Corner = @(c) reshape(abs(D0(:, c)), L, K);
            i11 = Corner(1);  iE1 = Corner(L);  i1E = Corner(LK - L + 1);  iEE = Corner(LK);
            Pattern = [[iEE; i1E(2:end, :)], [iE1(:, 2:end); i11(2:end, 2:end)]];
            Sorted = sort(Pattern(:), 'descend');
            Threshold = Sorted(NrCancel + 1);
            Mask = abs(D0(pil, :)) >= Threshold * (1 - obj.TieTolerance);
            Tags = -sum(bsxfun(@times, Mask, (1:numel(pil)).'), 1).';
            Tags(pil) = (1:numel(pil)).';
        end
    end
    methods (Static)
        function Out = SpreadingCode(Row)
            % orthonormal code over the interferers of one pilot whose weighted sum (weights = imaginary interference)
            % vanishes: Hadamard codes inside clusters of equal |interference|, pairwise links between the clusters
            % (smallest two merged first), then Gram-Schmidt
            w = imag(Row(:));
            w = round(abs(w) * 1e10) .* sign(w) / 1e10;
            n = numel(w);
            [~, Order] = sort(abs(w), 'descend');
            ws = w(Order);  Mags = abs(ws);
            Levels = unique(Mags);
            B = zeros(n, n - 1);
            Col = 0;
            Groups = {};
            for u = Levels(:).'
                idx = find(Mags == u);
                Groups{end + 1} = double(ismember((1:n).', idx)); %#ok<AGROW>
                m = numel(idx);
                if m > 1 && bitand(m, m - 1) == 0
                    H = 1;
                    while size(H, 1) < m, H = [H, H; H, -H]; end %#ok<AGROW>
                    Blk = bsxfun(@rdivide, H, ws(idx));
                    Blk = Blk(:, 2:end);
                elseif m > 1
                    E = eye(m, m - 1);
                    Blk = bsxfun(@rdivide, E - circshift(E, [1 0]), ws(idx));
                else
                    continue;
                end
                B(idx, Col + (1:size(Blk, 2))) = Blk;
                Col = Col + size(Blk, 2);
            end
            for i = 1:numel(Levels) - 1 %#ok<NASGU>
                [~, ia] = min(cellfun(@sum, Groups));  a = Groups{ia};  Groups(ia) = [];
                [~, ib] = min(cellfun(@sum, Groups));  b = Groups{ib};  Groups(ib) = [];
                ja = find(a, 1);  jb = find(b, 1);
                Col = Col + 1;
                B([ja jb], Col) = [1; -1] ./ ws([ja jb]);
                Groups{end + 1} = a + b; %#ok<AGROW>
            end
            Q = zeros(size(B));
            for c = 1:n - 1
                v = B(:, c) - Q(:, 1:c - 1) * (Q(:, 1:c - 1).' * B(:, c));
                Q(:, c) = v / sqrt(v.' * v);
            end
            Out = zeros(size(Q));
            Out(Order, :) = Q;
        end
    end
end
