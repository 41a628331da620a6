function worst = verify_oracle(reference_dir, bundle_file)
% VERIFY_ORACLE  Pin the repository's CPU oracle to the UNMODIFIED reference (MATLAB or GNU Octave).
%
%   verify_oracle('/path/to/Channel-Estimation', 'tests/golden/reference_bundle_default.mat')
%
% Runs the text of DoublySelectiveChannelEstimation.m with only its parameter block overridden (NrRepetitions,
% M_SNR_dB, optionally the paper block DS.m:42-46) while rand / randn / randi are shadowed by matlab/replay/,
% which pop the draws exported by oracle/export_reference_bundle.py in the order the script consumes them.
% Every class (+Channel, +Modulation, +ChannelEstimation) is the reference's own.  Afterwards the 24 BER arrays
% (DS.m:322-345) and the end-of-script workspace variables are compared with the oracle's values in the bundle.
% Expected: BER arrays identical (hard decisions), complex vectors within 1e-9 relative.
here = fileparts(mfilename('fullpath'));
addpath(fullfile(here, 'replay'));                      % shims first on the path
addpath(reference_dir);
chest_replay_queue('load', bundle_file);
ov = getfield(chest_replay_queue('bundle'), 'overrides');
txt = fileread(fullfile(reference_dir, 'DoublySelectiveChannelEstimation.m'));
txt = regexprep(txt, 'NrRepetitions\s*=\s*25;', sprintf('NrRepetitions = %d;', ov.NrRepetitions), 'once');
txt = regexprep(txt, 'M_SNR_dB\s*=\s*\[10:5:40\];', ['M_SNR_dB = [' sprintf('%g ', ov.M_SNR_dB) '];'], 'once');
txt = regexprep(txt, 'PlotIterationStepsSNRdB\s*=\s*35;', sprintf('PlotIterationStepsSNRdB = %g;', ov.M_SNR_dB(end)), 'once');
if ov.paper_block                                       % DS.m:42-46: un-comment, then re-apply the overrides
    txt = regexprep(txt, '%\s*(M_SNR_dB\s*=\s*\[10:2:40\];)', ['M_SNR_dB = [' sprintf('%g ', ov.M_SNR_dB) '];']);
    txt = regexprep(txt, '%\s*NrRepetitions\s*=\s*1000;', sprintf('NrRepetitions = %d;', ov.NrRepetitions));
    txt = regexprep(txt, '%\s*(SamplingRate\s*=\s*F\*14\*14;)', '$1');
    txt = regexprep(txt, '%\s*(NrSubframes\s*=\s*2;)', '$1');
end
eval(txt);                                              % the script clears the workspace itself (DS.m:12)
B = chest_replay_queue('bundle');                       % our own variables did not survive that clear
fprintf('streams left unread [rand randi randn]: %s (expected 0 0 0)\n', mat2str(chest_replay_queue('left')));
worst = 0;
names = fieldnames(B.ber);
for k = 1:numel(names)
    ref = B.ber.(names{k});
    got = eval(names{k});
    d = max(abs(got(:) - ref(:)));
    fprintf('%-58s max |diff| = %.3g\n', names{k}, d);
    worst = max(worst, d);
end
names = fieldnames(B.last);
for k = 1:numel(names)
    ref = B.last.(names{k});
    if strcmp(names{k}, 'ImpulseResponse'), got = ChannelModel.ImpulseResponse(:,:,1,1); else, got = eval(names{k}); end
    d = max(abs(got(:) - ref(:))) / max(abs(ref(:)));
    fprintf('%-58s rel. dev  = %.3g\n', names{k}, d);
    worst = max(worst, d);
end
S = B.setup;
chk = {'Kappa_Aux', Kappa_Aux; 'Kappa_Cod', Kappa_Cod; 'Kappa_OFDM', Kappa_OFDM; ...
       'DataPowerReduction_Aux', AuxiliaryMethod.DataPowerReduction; 'DataPowerReduction_Cod', CodingMethod.DataPowerReduction; ...
       'SIR_dB_Aux', AuxiliaryMethod.SIR_dB; 'SIR_dB_Cod', CodingMethod.SIR_dB; ...
       'ConsideredInterference_Aux', AuxiliaryMethod.ConsideredInterferenceMatrix; ...
       'ConsideredInterference_Cod', CodingMethod.ConsideredInterferenceMatrix; ...
       'nnz_PrecodingMatrix_Aux', nnz(AuxiliaryMethod.PrecodingMatrix); 'nnz_PrecodingMatrix_Cod', nnz(CodingMethod.PrecodingMatrix); ...
       'R_hP_FBMC', R_hP_FBMC; 'R_hP_OFDM', R_hP_OFDM};
for k = 1:size(chk, 1)
    ref = S.(chk{k, 1}); got = chk{k, 2};
    d = max(abs(double(got(:)) - ref(:))) / max(1e-300, max(abs(ref(:))));
    fprintf('setup  %-51s rel. dev  = %.3g\n', chk{k, 1}, d);
    worst = max(worst, d);
end
fprintf(['\nworst deviation %.3g.  ConsideredInterference_* shows which members of the interference tie group the\n' ...
         'reference''s floating-point >= picked (IIC.m:72-73,113-114; DESIGN.md section 2): if it differs, feed the\n' ...
         'reference''s own PrecodingMatrix to the library (chest_set_scheme takes it as an input).\n'], worst);
end
