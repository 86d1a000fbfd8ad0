function [J0, J1] = NPICost(newcases, inputs, weights)
% Drop-in for Tools/NPICost.m:1.
[J0, J1] = epi_mex('npicost', newcases(:)', inputs, weights .* ones(size(inputs)));
end
