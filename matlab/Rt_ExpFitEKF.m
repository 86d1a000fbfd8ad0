function [S_MINUS, S_PLUS, P_MINUS, P_PLUS, K_GAIN, S_SMOOTH, P_SMOOTH, innovations, rho] = Rt_ExpFitEKF(x, s_init, params, w_bar, v_bar, Ps_init, Q_w, R_v, beta, gamma, inv_monitor_len, order)
% Drop-in for Tools/Rt_ExpFitEKF.m:1 (2-state exponential-fit EKF/EKS, order 1 or 2).
if(order ~= 1 && order ~= 2), error('Undefined order'); end
[S_MINUS, S_PLUS, P_MINUS, P_PLUS, K_GAIN, S_SMOOTH, P_SMOOTH, innovations, rho] = ...
    epi_mex('rt_expfit', x, s_init(:), params(:), w_bar(:), v_bar, Ps_init, Q_w, R_v, beta, gamma, inv_monitor_len, order);
end
