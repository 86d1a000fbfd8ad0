function [s, e, i, r, p] = SEIRPSaturatedResource(alpha_e, alpha_i, kappa, rho, gamma, s0, e0, i0, r0, p0, T, dt, beta_0, beta_s, mu_0, mu_s, sigma, i_0)
% Drop-in for Tools/SEIRPSaturatedResource.m:1 -- runs on the B200 through libepi_b200 (epi_mex).
K = round(T/dt);
z = zeros(K, 1);
rates = [col(alpha_e, K), col(alpha_i, K), col(kappa, K), col(rho, K), z, z, col(gamma, K)]; % beta, mu rows are ignored
[s, e, i, r, p] = epi_mex('seirp', 1, rates, [s0; e0; i0; r0; p0], K, dt, [beta_0; beta_s; mu_0; mu_s; sigma; i_0]);
end

function c = col(v, K)
% K x 1 column; only samples 1..K-1 are read by the integrator (SEIRP.m:26)
if(numel(v) == 1)
    c = v * ones(K, 1);
else
    c = zeros(K, 1);
    n = min(K, numel(v));
    c(1:n) = v(1:n);
end
end
