function r = randn(varargin)
% Shim used only by matlab/verify_oracle.m: pops exported normals instead of drawing (DS.m:399).
r = chest_replay_queue('randn', chest_replay_size(varargin{:}));
end
