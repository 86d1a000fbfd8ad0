function [u_opt, u_opt_smooth, S_MINUS, S_PLUS, S_SMOOTH, P_MINUS, P_PLUS, P_SMOOTH, K_GAIN, innovations, rho] = SIAlphaModelEKFOptControlled(u, x, params, s_init, Ps_init, s_final, Ps_final, w_bar, v_bar, Q_w, R_v, beta, gamma, inv_monitor_len, order)
% Drop-in for Tools/SIAlphaModelEKFOptControlled.m:1 -- EPI_MODEL id 2 of include/epi_b200.h.
[u_opt, u_opt_smooth, S_MINUS, S_PLUS, S_SMOOTH, P_MINUS, P_PLUS, P_SMOOTH, K_GAIN, innovations, rho] = epi_mex('ekf_eks', 2, u, x(:)', params, s_init(:), Ps_init, s_final(:), Ps_final, w_bar, v_bar, Q_w, R_v, beta, gamma, inv_monitor_len, order);
end
