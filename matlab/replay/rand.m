function r = rand(varargin)
% Shim used only by matlab/verify_oracle.m: pops exported uniforms instead of drawing (FF.m:227-233).
r = chest_replay_queue('rand', chest_replay_size(varargin{:}));
end
