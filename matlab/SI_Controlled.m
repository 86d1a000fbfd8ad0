function [s, i] = SI_Controlled(alpha, beta, s0, i0, K, dt)
% Drop-in for Tools/SI_Controlled.m:1.
[s, i] = epi_mex('si_controlled', alpha(1:K), beta, s0, i0, K, dt);
end
