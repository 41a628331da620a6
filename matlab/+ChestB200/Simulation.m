classdef Simulation < handle
    % Tier-2 replacement of the loop body DoublySelectiveChannelEstimation.m:350-565 (and, optionally, of the MMSE
    % setup of lines 257-313) on one or several B200s driven from MATLAB's single interpreter thread.
    %
    % Run the reference script's parameter block and setup unchanged, then
    %
    %   sim = ChestB200.Simulation(ChannelModel, G_FBMC, Q_FBMC, G_OFDM, Q_OFDM, PAM, QAM, M_SNR_dB, ...
    %                              SamplingRate / (F * L), schemes, MaxBatch, NrGPUs);
    %   BER = sim.Run(NrRepetitions, NrIterations, seed);
    %
    % schemes(k) fields: id (0 Aux, 1 Cod, 2 OFDM), waveform (0 FBMC, 1 OFDM), C (PrecodingMatrix or
    % PilotMapping_OFDM), PilotMatrix, DataPositions ([] for Cod), Kappa, DataPowerReduction,
    % DetectMode (0/1/2), Constellation (0 PAM, 1 QAM), ConsideredBits, and EITHER
    %   W_MMSE, W_MMSE_noInterference      the sparse K^2 P x S matrices of DS.m:279-313 (host setup, lines 208-313 kept), OR
    %   R_hP_est_noNoise                   the P x P matrix of DS.m:213-234 only: lines 257-313 are then skipped and
    %                                      R_Dij_hP and W are built on the device (chest_setup_correlations /
    %                                      chest_build_mmse; ZeroThresholdSparse as in DS.m:20).
    % BER.<name> reproduces the 24 arrays of lines 322-345 (S x reps [x I]); the plotting code consumes them unchanged.
    % With NrGPUs > 1 one context per device is configured identically and chest_multi_run shards the realizations
    % (contiguous blocks, no inter-GPU traffic in the loop, one NCCL all-reduce of the counters at the end).
    % NOT EXECUTED in this repository's CI (no MATLAB / Octave in the image); see INTEGRATION.md.
    properties (SetAccess = private)
        Handles            % uint64, one context per device
        Multi = []         % chest_multi handle when NrGPUs > 1
        NrSNR
        Schemes
        MaxBatch
        LastReduceMs = 0
    end
    methods
        function obj = Simulation(ChannelModel, G_F, Q_F, G_O, Q_O, PAM, QAM, M_SNR_dB, NoiseFactor, schemes, MaxBatch, NrGPUs, ZeroThresholdSparse)
            if nargin < 11 || isempty(MaxBatch), MaxBatch = 1024; end
            if nargin < 12 || isempty(NrGPUs), NrGPUs = 1; end
            if nargin < 13, ZeroThresholdSparse = 8; end
            obj.NrSNR = numel(M_SNR_dB);  obj.Schemes = schemes;  obj.MaxBatch = MaxBatch;
            Pn = NoiseFactor * 10.^(-M_SNR_dB(:) / 10);
            model = find(strcmp(ChannelModel.PHY.DopplerModel, {'Jakes', 'Uniform', 'Discrete-Jakes', 'Discrete-Uniform'})) - 1;
            obj.Handles = zeros(1, NrGPUs, 'uint64');
            for dev = 1:NrGPUs
                h = chest_mex('create', dev - 1);
                obj.Handles(dev) = h;
                chest_mex('set_channel', h, ChannelModel.Nr.SamplesTotal, ChannelModel.Implementation.PowerDelayProfileNormalized, ...
                    ChannelModel.PHY.MaximumDopplerShift, ChannelModel.PHY.dt, ChannelModel.Nr.Paths, model);
                chest_mex('set_waveform', h, 0, G_F, Q_F);
                chest_mex('set_waveform', h, 1, G_O, Q_O);
                chest_mex('set_constellation', h, 0, PAM.SymbolMapping, double(PAM.BitMapping));
                chest_mex('set_constellation', h, 1, QAM.SymbolMapping, double(QAM.BitMapping));
                chest_mex('set_snr', h, Pn);
                OnDevice = isfield(schemes, 'R_hP_est_noNoise') && ~isempty(schemes(1).R_hP_est_noNoise);
                if OnDevice
                    % the correlation pass runs the P pseudo-channels through K1 + K2: it needs a finalized context
                    P = max(arrayfun(@(s) nnz(s.PilotMatrix == 1), schemes));
                    chest_mex('finalize', h, P);
                end
                for s = schemes(:).'
                    chest_mex('set_scheme', h, s.id, s.waveform, sparse(s.C), find(s.PilotMatrix(:) == 1), s.DataPositions, ...
                        s.Kappa, s.DataPowerReduction, s.DetectMode, s.Constellation, double(s.ConsideredBits));
                end
                if OnDevice
                    obj.EstimatorSetupOnDevice(h, ChannelModel, {Q_F, Q_O}, Pn, 10^(-ZeroThresholdSparse));
                else
                    for s = schemes(:).'
                        chest_mex('set_mmse', h, s.id, 0, s.W_MMSE);
                        chest_mex('set_mmse', h, s.id, 1, s.W_MMSE_noInterference);
                    end
                end
                chest_mex('finalize', h, MaxBatch);
            end
            if NrGPUs > 1, obj.Multi = chest_mex('multi_create', obj.Handles); end
        end

        function BER = Run(obj, NrRepetitions, NrIterations, seed, FirstRepetition)
            % seeded realizations (counter-based generator keyed by (seed, realization index): the result does not
            % depend on MaxBatch or on the number of GPUs)
            if nargin < 5, FirstRepetition = 0; end
            if ~isempty(obj.Multi)
                [err, ~, obj.LastReduceMs] = chest_mex('multi_run', obj.Multi, NrRepetitions, NrIterations, seed, FirstRepetition, obj.NrSNR);
            else
                err = zeros(12 * (NrIterations + 1) * obj.NrSNR, NrRepetitions, 'uint32');
                for r0 = 0:obj.MaxBatch:NrRepetitions - 1
                    n = min(obj.MaxBatch, NrRepetitions - r0);
                    err(:, r0 + (1:n)) = chest_mex('run_batch', obj.Handles(1), n, NrIterations, seed, FirstRepetition + r0, obj.NrSNR);
                end
            end
            BER = obj.BerArrays(err, NrIterations);
        end

        function SetEstimatorMode(obj, Mode)
            % Form of the estimated-CSI cancellation (include/chest_b200.h): 'auto' (default), 'tiles' (the thresholded W_MMSE of
            % DS.m:279-313, bit-faithful), 'factored' (stated-tolerance mode: D_est = Q' H_est G with the estimated channel
            % H_est = sum_q g_q M_q, without the two 1e-8 thresholds), 'factored_exact' (factored only where the thresholds
            % removed nothing, i.e. CP-OFDM).  Needs the device-side setup (R_hP_est_noNoise given) and set_modem descriptions.
            m = find(strcmp(Mode, {'auto', 'tiles', 'factored', 'factored_exact'})) - 1;
            assert(~isempty(m), 'Mode must be auto, tiles, factored or factored_exact');
            for h = obj.Handles, chest_mex('set_estimator_mode', h, m); end
        end

        function BER = RunWithDraws(obj, NrIterations, Draws)
            % explicit draws exported from a MATLAB run (bit-exact replay of DS.m:352-368,399); Draws fields, one column
            % (page for Noise) per realization: DopplerU, PhaseU (T*Paths x reps, rand([T 1 Paths]) order), BitsAux, BitsCod,
            % BitsOFDM, PilotIndexFBMC, PilotIndexOFDM (1-based SymbolMapping indices), Noise (N x S x reps complex)
            NrRepetitions = size(Draws.DopplerU, 2);
            err = zeros(12 * (NrIterations + 1) * obj.NrSNR, NrRepetitions, 'uint32');
            for r0 = 0:obj.MaxBatch:NrRepetitions - 1
                c = r0 + (1:min(obj.MaxBatch, NrRepetitions - r0));
                err(:, c) = chest_mex('run_batch_draws', obj.Handles(1), numel(c), NrIterations, obj.NrSNR, ...
                    Draws.DopplerU(:, c), Draws.PhaseU(:, c), Draws.BitsAux(:, c), Draws.BitsCod(:, c), Draws.BitsOFDM(:, c), ...
                    Draws.PilotIndexFBMC(:, c), Draws.PilotIndexOFDM(:, c), Draws.Noise(:, :, c));
            end
            BER = obj.BerArrays(err, NrIterations);
        end

        function delete(obj)
            if ~isempty(obj.Multi), chest_mex('multi_destroy', obj.Multi); obj.Multi = []; end
            for h = obj.Handles, chest_mex('destroy', h); end
            obj.Handles = [];
        end
    end
    methods (Access = private)
        function EstimatorSetupOnDevice(obj, h, ChannelModel, Q, Pn, Threshold)
            % DS.m:257-313 on the device.  R_hP (DS.m:213) comes back from the correlation pass; the noise / interference
            % terms of R_hP_est (DS.m:238-253) and the P x P pseudo-inverses (DS.m:283-285) are P x P host work.
            Rt = ChannelModel.GetTimeCorrelation;
            Done = [false false];  R_hP = cell(1, 2);
            for s = obj.Schemes(:).'
                w = s.waveform + 1;
                pil = find(s.PilotMatrix(:) == 1);  P = numel(pil);
                if ~Done(w)
                    R_hP{w} = chest_mex('setup_correlations', h, s.waveform, pil, Rt, Threshold);
                    Done(w) = true;
                end
            end
            for s = obj.Schemes(:).'
                w = s.waveform + 1;
                pil = find(s.PilotMatrix(:) == 1);  P = numel(pil);
                qn = real(sum(abs(Q{w}(:, pil)).^2, 1)).';
                for variant = 0:1
                    Rinv = zeros(P, P, numel(Pn));
                    for i = 1:numel(Pn)
                        R = s.R_hP_est_noNoise;
                        R(1:P + 1:end) = diag(s.R_hP_est_noNoise) + Pn(i) * qn / s.Kappa;
                        if variant == 1, R = R - (s.R_hP_est_noNoise - R_hP{w}); end
                        Rinv(:, :, i) = pinv(R);
                    end
                    chest_mex('build_mmse', h, s.id, variant, Rinv, Threshold);
                end
            end
            chest_mex('release_setup', h);
        end

        function BER = BerArrays(obj, err, NrIterations)
            % err(edge, csi, scheme, it, snr, rep): column-major view of the ABI's [rep][snr][it][scheme][csi][edge]
            NrRepetitions = size(err, 2);
            err = double(reshape(err, 2, 2, 3, NrIterations + 1, obj.NrSNR, NrRepetitions));
            nb = chest_mex('bit_counts', obj.Handles(1));      % 2 x 3
            names = {'FBMC_Aux', 'FBMC_Cod', 'OFDM'};  csi = {'', '_PerfectCSI'};  edge = {'', '_NoEdge'};
            for s = obj.Schemes(:).'
                for c = 1:2
                    for e = 1:2
                        x = permute(reshape(err(e, c, s.id + 1, :, :, :), NrIterations + 1, obj.NrSNR, NrRepetitions), [2 3 1]) / nb(e, s.id + 1);   % S x reps x (1+I)
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
    end
end
