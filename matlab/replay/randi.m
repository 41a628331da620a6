function r = randi(imax, varargin)
% Shim used only by matlab/verify_oracle.m: pops exported integers instead of drawing (DS.m:355-367).
% The range argument is only checked: randi([0 1], n, 1) -> values in {0,1}; randi(M, n, 1) -> 1..M.
r = chest_replay_queue('randi', chest_replay_size(varargin{:}));
lo = 1; hi = imax(end);
if numel(imax) == 2, lo = imax(1); end
if any(r(:) < lo) || any(r(:) > hi)
    error('chest:replay', 'randi replay out of the requested range [%g, %g]: stream order mismatch', lo, hi);
end
end
