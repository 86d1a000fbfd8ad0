function [u_opt, S_MINUS, S_PLUS, S_SMOOTH, P_MINUS, P_PLUS, P_SMOOTH, K_GAIN, innovations, rho] = NewCaseEKFEstimatorWithOptimalNPI(u, x, params, s_init, Ps_init, s_final, Ps_final, w_bar, v_bar, Q_w, R_v, beta, gamma, inv_monitor_len, order)
% Drop-in for Tools/NewCaseEKFEstimatorWithOptimalNPI.m:1 (output order of the Tools/ copy;
% EPI_MODEL_LEGACY_TOOLS = 4).  For the MatlabCodeGenerator/ copy (different output order,
% identity ObsHardMargins) use model id 5 and permute the outputs as in
% MatlabCodeGenerator/NewCaseEKFEstimatorWithOptimalNPI.m:1.
[u_opt, ~, S_MINUS, S_PLUS, S_SMOOTH, P_MINUS, P_PLUS, P_SMOOTH, K_GAIN, innovations, rho] = epi_mex('ekf_eks', 4, u, x(:)', params, s_init(:), Ps_init, s_final(:), Ps_final, w_bar, v_bar, Q_w, R_v, beta, gamma, inv_monitor_len, order);
end
