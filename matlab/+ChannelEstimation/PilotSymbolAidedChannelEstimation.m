classdef PilotSymbolAidedChannelEstimation < handle
    % Drop-in for the reference's ChannelEstimation.PilotSymbolAidedChannelEstimation
    % (+ChannelEstimation/PilotSymbolAidedChannelEstimation.m:33-184): pilot patterns ('Rectangular', 'Diamond',
    % 'Custom') and the interpolation of the LS estimates over the L x K grid ('linear' / 'nearest' / 'natural' through
    % MATLAB's scatteredInterpolant exactly as the reference does, 'FullAverage', 'MovingBlockAverage').
    % GetInterpolationMatrix is what the device path consumes (chest_mex('set_interpolation', ...), SV.m:143-145).
    % NOT EXECUTED in this repository's CI (no MATLAB / Octave); mirror of chest_b200.ChannelEstimation.*.
    properties (SetAccess = private)
        NrPilotSymbols
        PilotPattern
        PilotSpacingFrequency
        PilotSpacingTime
        InterpolationMethod
        Implementation
        InterpolationProperties
        PilotMatrix
    end
    methods
        function obj = PilotSymbolAidedChannelEstimation(PilotPattern, PatternParameters, InterpolationMethod, BlockLengths)
            obj.PilotPattern = PilotPattern;
            obj.InterpolationMethod = InterpolationMethod;
            switch PilotPattern
                case {'Rectangular', 'Diamond'}
                    nL = PatternParameters(1, 1);  dF = PatternParameters(1, 2);
                    nK = PatternParameters(2, 1);  dT = PatternParameters(2, 2);
                    obj.PilotSpacingFrequency = dF;  obj.PilotSpacingTime = dT;
                    pm = zeros(nL, nK);
                    if strcmp(PilotPattern, 'Rectangular')
                        f0 = round(mod(nL - 1, dF) / 2) + 1;
                        t0 = round(round(mod(nK - 1, dT) / 2)) + 1;
                        pm(f0:dF:nL, t0:dT:nK) = 1;
                    else
                        Last = @(s, st, n) max([s:st:n, -inf]);
                        fmax = max([Last(1, 2 * dF, nL), Last(1 + dF / 2, 2 * dF, nL), Last(1 + dF, 2 * dF, nL), Last(1 + 3 * dF / 2, 2 * dF, nL)]);
                        tmax = max([Last(1, 2 * dT, nK), Last(1 + dT, 2 * dT, nK)]);
                        fs = floor((nL - fmax) / 2) + 1;
                        ts = floor((nK - tmax) / 2) + 1;
                        pm(fs:2 * dF:nL, ts:2 * dT:nK) = 1;
                        pm(fs + round(dF / 2):2 * dF:nL, round(ts + dT):2 * dT:nK) = 1;
                        pm(fs + round(dF):2 * dF:nL, ts:2 * dT:nK) = 1;
                        pm(fs + round(3 * dF / 2):2 * dF:nL, round(ts + dT):2 * dT:nK) = 1;
                    end
                case 'Custom'
                    obj.PilotSpacingFrequency = nan;  obj.PilotSpacingTime = nan;
                    pm = PatternParameters;
                otherwise
                    error('Pilot pattern is not supported! Chose Rectangular Diamond or Custom');
            end
            obj.PilotMatrix = pm;
            obj.NrPilotSymbols = sum(pm(:));
            switch InterpolationMethod
                case {'linear', 'nearest', 'natural'}
                    [f, t] = find(pm);
                    obj.InterpolationProperties = scatteredInterpolant(f, t, zeros(numel(f), 1), InterpolationMethod);
                case 'MovingBlockAverage'
                    bF = BlockLengths(1);  bT = BlockLengths(2);
                    [nL, nK] = size(pm);
                    Numbered = zeros(nL, nK);
                    Numbered(pm == 1) = 1:obj.NrPilotSymbols;
                    M = zeros(nL * nK, obj.NrPilotSymbols);
                    for Pos = 1:nL * nK
                        [f, t] = ind2sub([nL nK], Pos);
                        Blk = Numbered(max(f - bF, 1):min(f + bF, nL), max(t - bT, 1):min(t + bT, nK));
                        Sel = Blk(Blk > 0);
                        M(Pos, Sel) = 1 / numel(Sel);
                    end
                    obj.InterpolationProperties.InterpolationMatrix = M;
                case 'FullAverage'
                case 'MMSE'
                    error('Needs to be implemented');
                otherwise
                    error('Interpolation method not implemented');
            end
        end

        function InterpolatedChannel = ChannelInterpolation(obj, LSChannelEstimatesAtPilotPosition)
            pm = obj.PilotMatrix;
            switch obj.InterpolationMethod
                case 'FullAverage'
                    InterpolatedChannel = ones(size(pm)) * mean(LSChannelEstimatesAtPilotPosition(:));
                case 'MovingBlockAverage'
                    InterpolatedChannel = reshape(obj.InterpolationProperties.InterpolationMatrix * LSChannelEstimatesAtPilotPosition(:), size(pm));
                otherwise
                    obj.InterpolationProperties.Values = LSChannelEstimatesAtPilotPosition(:);
                    [f, t] = ndgrid(1:size(pm, 1), 1:size(pm, 2));
                    InterpolatedChannel = obj.InterpolationProperties(f, t);
            end
        end

        function InterpolationMatrix = GetInterpolationMatrix(obj)
            % column i = interpolated response to a unit estimate at pilot i: the interpolators are linear in the
            % pilot estimates, so h_interp(:) = InterpolationMatrix * h_LS
            P = obj.NrPilotSymbols;
            InterpolationMatrix = zeros(numel(obj.PilotMatrix), P);
            for i = 1:P
                e = zeros(P, 1);  e(i) = 1;
                x = obj.ChannelInterpolation(e);
                InterpolationMatrix(:, i) = x(:);
            end
        end

        function AuxiliaryMatrix = GetAuxiliaryMatrix(obj, NrAxuiliarySymbols)
            % pilot matrix with -1 at the auxiliary positions next to every pilot (time neighbours first)
            if ~any(NrAxuiliarySymbols == [1 2 3 4]), error('Only 1,2,3,4 auxiliary symbols per pilot are supported'); end
            AuxiliaryMatrix = obj.PilotMatrix;
            [ls, ks] = find(obj.PilotMatrix);
            for i = 1:numel(ls)
                l = ls(i);  k = ks(i);
                AuxiliaryMatrix(l, k + 1) = -1;
                if NrAxuiliarySymbols >= 2, AuxiliaryMatrix(l, k - 1) = -1; end
                if NrAxuiliarySymbols >= 3, AuxiliaryMatrix(l + 1, k) = -1; end
                if NrAxuiliarySymbols >= 4, AuxiliaryMatrix(l - 1, k) = -1; end
            end
        end
    end
end
