function [s, i, alpha] = SIalpha_Controlled(u, s0, i0, alpha0, u_max, alpha_min, alpha_max, gamma, a, b, beta, s_noise_std, i_noise_std, alpha_noise_std, K, dt)
% Drop-in for Tools/SIalpha_Controlled.m:1.  The reference draws randn in-line
% (:25-27, call order s, i, alpha per step); the same draws are made here on the
% host, in the same order, and handed to the device as an explicit input.
params = struct('dt', dt, 'beta', beta, 'gamma', gamma, 'b', b, 'alpha_min', alpha_min, 'alpha_max', alpha_max, 'a', a(:), 'u_max', u_max(:));
noise = randn(3, K); % column k = the three draws of step k, in the reference's order
[s, i, alpha] = epi_mex('sialpha_controlled', u(:, 1:K), [s0; i0; alpha0], params, [s_noise_std; i_noise_std; alpha_noise_std], K, noise);
end
