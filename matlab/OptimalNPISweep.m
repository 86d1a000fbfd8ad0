function [J0, J1, on_front, I_opt, u_knee] = OptimalNPISweep(params, epsilons, u, x, R_v, s_init, Ps_init, s_final, Ps_final, Q_w, beta_ekf, gamma_ekf, inv_monitor_len, x0, newcases_hist, weights, lean, n_gpus)
% Batched replacement of the epsilon loop of Tools/TrainPredictPrescribeNPI.m:421-495 and the Pareto
% step :624-633 for ALL regions at once (one call into libepi_b200's epi_sweep):
%   params        1 x nR struct array (the reference's params with .w = NPI cost weights; .epsilon ignored)
%   epsilons      1 x nE human/NPI cost trade-off grid (human_npi_cost_factor)
%   u             L x T x nR control_input = [IP, nan(L, num_forecast_days)] per region (:458)
%   x, R_v        T x nR observations (NaN on forecast days, :364) and their variances (:360)
%   s_init, Ps_init, s_final, Ps_final, Q_w   6 x nR / 6 x 6 x nR  (ss_init, PPs_init, ..., QQ_w of :428-457)
%   x0            3 x nR rollout start [s_historic(end); i_historic(end); alpha_historic(end)] (:481)
%   newcases_hist T_hist x nR  s_historic.*i_historic.*alpha_historic (:493)
%   weights       L x T x nR  npi_weights_day_wise (:390)
%   lean          1 = smooth only the days to optimise (identical outputs, faster); 0 = all T days
%   n_gpus        GPUs of the box to shard the regions over in this one blocking call (default 1; 0 = all):
%                 the region loop of Tools/TrainPredictPrescribeNPI.m:93 spread over the devices, same outputs
% Returns J0_opt_control / J1_opt_control (nE x nR), the Pareto mask, the knee index (1-based) and the
% prescribed schedule at the knee (L x T_fore x nR).
if(nargin < 17), lean = 1; end
if(nargin < 18), n_gpus = 1; end
[J0, J1, on_front, I_opt, u_knee] = epi_mex('sweep', params, epsilons(:)', u, x, R_v, s_init, Ps_init, s_final, Ps_final, Q_w, beta_ekf, gamma_ekf, inv_monitor_len, x0, newcases_hist, weights, lean, n_gpus);
end
