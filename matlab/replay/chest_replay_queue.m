function out = chest_replay_queue(kind, varargin)
% Replay queues behind the rand / randn / randi shims of this directory (matlab/verify_oracle.m).
%   chest_replay_queue('load', bundle_file)   load the three streams exported by oracle/export_reference_bundle.py
%   chest_replay_queue('rand', sz)            next prod(sz) values of that stream, reshaped to sz
%   chest_replay_queue('bundle')              the loaded bundle (survives the script's own `clear`)
%   chest_replay_queue('left')                values left per stream [rand randi randn]
persistent B pos
switch kind
    case 'load'
        B = load(varargin{1});
        pos = struct('rand', 0, 'randi', 0, 'randn', 0);
        out = [];
    case 'bundle'
        out = B;
    case 'left'
        out = [numel(B.stream_rand) - pos.rand, numel(B.stream_randi) - pos.randi, numel(B.stream_randn) - pos.randn];
    otherwise
        sz = varargin{1};
        n = prod(sz);
        s = B.(['stream_' kind]);
        if pos.(kind) + n > numel(s)
            error('chest:replay', 'replay stream "%s" exhausted', kind);
        end
        out = reshape(s(pos.(kind) + (1:n)), sz);
        pos.(kind) = pos.(kind) + n;
end
end
