classdef Simulation < handle
    % Tier-2 replacement of the loop body DoublySelectiveChannelEstimation.m:350-565.
    % Run the reference script's own setup (lines 12-346) unchanged, then
    %
    %   sim = ChestB200.Simulation(ChannelModel, G_FBMC, Q_FBMC, G_OFDM, Q_OFDM, PAM, QAM, M_SNR_dB, ...
    %                              SamplingRate / (F * L), schemes);   % schemes: struct array, see below
    %   BER = sim.Run(NrRepetitions, NrIterations, seed);
    %
    % schemes(k) fields: id (0 Aux, 1 Cod, 2 OFDM), waveform (0 FBMC, 1 OFDM), C (PrecodingMatrix or
    % PilotMapping_OFDM), PilotMatrix, DataMatrix ([] for Cod), Kappa, DataPowerReduction,
    % DetectMode (0/1/2), Constellation (0 PAM, 1 QAM), ConsideredBits, W_MMSE, W_MMSE_noInterference.
    % BER.<name> reproduces the 24 arrays of lines 322-345.
    % NOT EXECUTED in this repository's CI (no MATLAB / Octave in the image); see INTEGRATION.md.
    properties (SetAccess = private)
        Handle
        NrSNR
        Schemes
    end
    methods
        function obj = Simulation(ChannelModel, G_F, Q_F, G_O, Q_O, PAM, QAM, M_SNR_dB, NoiseFactor, schemes, MaxBatch)
            if nargin < 11, MaxBatch = 1024; end
            h = chest_mex('create', 0);
            obj.Handle = h;  obj.NrSNR = numel(M_SNR_dB);  obj.Schemes = schemes;
            model = find(strcmp(ChannelModel.PHY.DopplerModel, {'Jakes', 'Uniform'})) - 1;
            chest_mex('set_channel', h, ChannelModel.Nr.SamplesTotal, ChannelModel.Implementation.PowerDelayProfileNormalized, ...
                ChannelModel.PHY.MaximumDopplerShift, ChannelModel.PHY.dt, ChannelModel.Nr.Paths, model);
            chest_mex('set_waveform', h, 0, G_F, Q_F);
            chest_mex('set_waveform', h, 1, G_O, Q_O);
            chest_mex('set_constellation', h, 0, PAM.SymbolMapping, double(PAM.BitMapping));
            chest_mex('set_constellation', h, 1, QAM.SymbolMapping, double(QAM.BitMapping));
            chest_mex('set_snr', h, NoiseFactor * 10.^(-M_SNR_dB(:) / 10));
            for s = schemes(:).'
                chest_mex('set_scheme', h, s.id, s.waveform, sparse(s.C), find(s.PilotMatrix(:) == 1), s.DataPositions, ...
                    s.Kappa, s.DataPowerReduction, s.DetectMode, s.Constellation, double(s.ConsideredBits));
                chest_mex('set_mmse', h, s.id, 0, s.W_MMSE);
                chest_mex('set_mmse', h, s.id, 1, s.W_MMSE_noInterference);
            end
            chest_mex('finalize', h, MaxBatch);
        end
        function BER = Run(obj, NrRepetitions, NrIterations, seed)
            err = chest_mex('run_batch', obj.Handle, NrRepetitions, NrIterations, seed, 0, obj.NrSNR);
            % err(edge, csi, scheme, it, snr, rep): column-major view of the ABI's [rep][snr][it][scheme][csi][edge]
            err = double(reshape(err, 2, 2, 3, NrIterations + 1, obj.NrSNR, NrRepetitions));
            nb = chest_mex('bit_counts', obj.Handle);      % 2 x 3
            names = {'FBMC_Aux', 'FBMC_Cod', 'OFDM'};  csi = {'', '_PerfectCSI'};  edge = {'', '_NoEdge'};
            for s = obj.Schemes(:).'
                for c = 1:2
                    for e = 1:2
                        x = permute(squeeze(err(e, c, s.id + 1, :, :, :)), [2 3 1]) / nb(e, s.id + 1);   % S x reps x (1+I)
                        BER.(['BER_' names{s.id + 1} '_OneTapEqualizer' csi{c} edge{e}]) = x(:, :, 1);
                        if c == 1
                            BER.(['BER_' names{s.id + 1} '_InterferenceCancellation' edge{e}]) = x(:, :, 2:end);
                        else
                            BER.(['BER_' names{s.id + 1} '_PerfectCSI_InterferenceCancellation' edge{e}]) = x(:, :, 2:end);
                        end
                    end
                end
            end
        end
        function delete(obj)
            if ~isempty(obj.Handle), chest_mex('destroy', obj.Handle); end
        end
    end
end
