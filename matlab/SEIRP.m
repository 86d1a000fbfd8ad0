function [s, e, i, r, p] = SEIRP(alpha_e, alpha_i, kappa, rho, beta, mu, gamma, s0, e0, i0, r0, p0, T, dt)
% Drop-in for Tools/SEIRP.m:1 -- runs on the B200 through libepi_b200 (epi_mex).
K = round(T/dt);
rates = [col(alpha_e, K), col(alpha_i, K), col(kappa, K), col(rho, K), col(beta, K), col(mu, K), col(gamma, K)]; % K x 7
[s, e, i, r, p] = epi_mex('seirp', 0, rates, [s0; e0; i0; r0; p0], K, dt, zeros(6, 1));
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
