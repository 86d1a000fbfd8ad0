function [EstError_PLUS, EstError_SMOOTH, S_PLUS_partial, S_SMOOTH_partial] = ForecastLookAheadErrors(control_input_ENTIRE, observations_ENTIRE, params, s_init, Ps_init, s_final, Ps_final, w_bar, v_bar, Q_w, R_v, beta_ekf, gamma_ekf, inv_monitor_len_ekf, num_forecast_days, MaxLookAheadDays, NewCasesSmoothed_ENTIRE, N_population)
% Batched replacement of the forecast look-ahead loop of Tools/ForecastQualityAssessment.m:383-394
% ("for start = 1 : num_forecast_days ... SIAlphaModelEKF(control_input_ENTIRE, observations_PARTIAL, ...)"):
% the num_forecast_days masked-horizon re-runs of the 3-state EKF/EKS are ONE call into libepi_b200
% (epi_mex('ekf_eks_masked', ...): one trajectory per `start`, the observations masked per trajectory), and
% the error tables are filled exactly as :387-394 does.  Variable names are the reference's.
%   EstError_PLUS, EstError_SMOOTH   num_forecast_days x MaxLookAheadDays (percent errors, zeros where the
%                                    reference leaves its preallocated zeros)
%   S_PLUS_partial, S_SMOOTH_partial 3 x T x num_forecast_days (page `start` = that iteration's estimates)
observations_ENTIRE = observations_ENTIRE(:)';
LL = length(observations_ENTIRE);
[S_PLUS_partial, S_SMOOTH_partial] = epi_mex('ekf_eks_masked', 0, control_input_ENTIRE, observations_ENTIRE, params, s_init(:), Ps_init, s_final(:), Ps_final, w_bar, v_bar, Q_w, R_v, beta_ekf, gamma_ekf, inv_monitor_len_ekf, 1, num_forecast_days);
EstError_PLUS = zeros(num_forecast_days, MaxLookAheadDays);
EstError_SMOOTH = zeros(num_forecast_days, MaxLookAheadDays);
truth = NewCasesSmoothed_ENTIRE(:)';
for start = 1 : num_forecast_days
    SP = S_PLUS_partial(:, :, start);
    SS = S_SMOOTH_partial(:, :, start);
    NewCasesSmoothed_EST_PLUS = N_population * SP(1, :) .* SP(2, :) .* SP(3, :);             % :387
    NewCasesSmoothed_EST_SMOOTH = N_population * SS(1, :) .* SS(2, :) .* SS(3, :);           % :388
    error_PLUS = 100 * abs(truth - NewCasesSmoothed_EST_PLUS) ./ truth;                      % :389
    error_SMOOTH = 100 * abs(truth - NewCasesSmoothed_EST_SMOOTH) ./ truth;                  % :390
    last_index = min(LL, LL - start + MaxLookAheadDays);                                     % :392
    EstError_PLUS(start, 1 : last_index - LL + start) = error_PLUS(LL - start + 1 : last_index);     % :393
    EstError_SMOOTH(start, 1 : last_index - LL + start) = error_SMOOTH(LL - start + 1 : last_index); % :394
end
end
