function sz = chest_replay_size(varargin)
% size arguments of rand/randn/randi: (n), (m,n,...), ([m n ...])
if nargin == 0
    sz = [1 1];
elseif nargin == 1
    sz = varargin{1};
    if isscalar(sz), sz = [sz sz]; end
else
    sz = [varargin{:}];
end
end
