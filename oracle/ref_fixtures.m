function ref_fixtures(ref_root, in_file, out_file)
% REF_FIXTURES  Run the UNMODIFIED reference (alphanumericslab/EpidemicModeling) on the parity cases.
%
%   ref_fixtures('/path/to/EpidemicModeling')
%
% Loads tests/golden/ref_inputs.mat (written by tools/make_ref_inputs.py from tests/cases.py), calls the
% reference's own Tools/*.m functions on every case and saves tests/golden/ref_outputs.mat.  With that file
% present, `pytest tests/test_oracle.py -k reference_fixtures` compares the C oracle (oracle/epi_oracle.c)
% with the reference itself -- the step that turns "parity unpinned" into a pinned oracle.  Works under
% MATLAB and GNU Octave; nothing of this repo's product code is on the path.  Test infrastructure.
%
% What is compared afterwards (tests/test_oracle.py): SEIRP, the rollout, NPICost and every forward pass
% (S_MINUS, S_PLUS, P_MINUS, P_PLUS, K_GAIN, innovations, rho) to rel 1e-9; the 3-state smoothers to 1e-9;
% the 6-state smoothers (cond(P_MINUS) up to 1e64, GenericExtendedKalmanFilter.m:215) by the number of
% bang-bang schedule entries that flip and by S_SMOOTH over the well-conditioned prefix.
here = fileparts(mfilename('fullpath'));
if nargin < 2 || isempty(in_file),  in_file  = fullfile(here, '..', 'tests', 'golden', 'ref_inputs.mat');  end
if nargin < 3 || isempty(out_file), out_file = fullfile(here, '..', 'tests', 'golden', 'ref_outputs.mat'); end
addpath(fullfile(ref_root, 'Tools'));
in = load(in_file);
out = struct();

% ---- SEIRP (Tools/SEIRP.m, Tools/SEIRPSaturatedResource.m)
names = fieldnames(in.seirp);
for k = 1 : numel(names)
    c = in.seirp.(names{k});
    [s, e, i, r, p] = SEIRP(c.alpha_e, c.alpha_i, c.kappa, c.rho, c.beta, c.mu, c.gamma, ...
                            c.s0, c.e0, c.i0, c.r0, c.p0, c.T, c.dt);
    out.seirp.(names{k}) = [s(:)'; e(:)'; i(:)'; r(:)'; p(:)'];
end
c = in.seirp_sat;
[s, e, i, r, p] = SEIRPSaturatedResource(c.alpha_e, c.alpha_i, c.kappa, c.rho, c.gamma, c.s0, c.e0, c.i0, ...
                                         c.r0, c.p0, c.T, c.dt, c.beta_0, c.beta_s, c.mu_0, c.mu_s, c.sigma, c.i_0);
out.seirp_sat = [s(:)'; e(:)'; i(:)'; r(:)'; p(:)'];

% ---- EKF / EKS, every model wrapper (GenericExtendedKalmanFilter.m underneath)
names = fieldnames(in.ekf);
for k = 1 : numel(names)
    c = in.ekf.(names{k});
    f = str2func(c.fn);
    o = struct();
    if strcmp(c.fn, 'NewCaseEKFEstimatorWithOptimalNPI')   % legacy monolith: no u_opt_smooth output
        [o.u_opt, o.S_MINUS, o.S_PLUS, o.S_SMOOTH, o.P_MINUS, o.P_PLUS, o.P_SMOOTH, o.K_GAIN, o.innovations, o.rho] = ...
            f(c.u, c.x, c.params, c.s_init, c.Ps_init, c.s_final, c.Ps_final, c.w_bar, c.v_bar, c.Q_w, c.R_v, ...
              c.beta, c.gamma, c.inv_monitor_len, c.order);
    else
        [o.u_opt, o.u_opt_smooth, o.S_MINUS, o.S_PLUS, o.S_SMOOTH, o.P_MINUS, o.P_PLUS, o.P_SMOOTH, o.K_GAIN, ...
         o.innovations, o.rho] = ...
            f(c.u, c.x, c.params, c.s_init, c.Ps_init, c.s_final, c.Ps_final, c.w_bar, c.v_bar, c.Q_w, c.R_v, ...
              c.beta, c.gamma, c.inv_monitor_len, c.order);
    end
    out.ekf.(names{k}) = o;
end

% ---- rollout + cost (Tools/SIalpha_Controlled.m with the recorded randn stream, Tools/NPICost.m)
c = in.rollout;
global EPI_REF_RANDN_STREAM EPI_REF_RANDN_POS
EPI_REF_RANDN_STREAM = c.randn_stream(:);
EPI_REF_RANDN_POS = 0;
addpath(fullfile(here, 'ref_shims'));          % shadows randn for this one call
[s, i, alpha] = SIalpha_Controlled(c.u, c.s0, c.i0, c.alpha0, c.u_max, c.alpha_min, c.alpha_max, c.gamma, ...
                                   c.a, c.b, c.beta, c.s_noise_std, c.i_noise_std, c.alpha_noise_std, c.K, c.dt);
rmpath(fullfile(here, 'ref_shims'));
assert(EPI_REF_RANDN_POS == numel(EPI_REF_RANDN_STREAM), 'randn stream not consumed exactly');
[J0, J1] = NPICost(s .* i .* alpha, c.u, c.weights);
out.rollout = struct('s', s, 'i', i, 'alpha', alpha, 'J', [J0, J1]);

% ---- Rt_ExpFitEKF (Tools/Rt_ExpFitEKF.m)
c = in.rt_expfit;
o = struct();
[o.S_MINUS, o.S_PLUS, o.P_MINUS, o.P_PLUS, o.K_GAIN, o.S_SMOOTH, o.P_SMOOTH, o.innovations, o.rho] = ...
    Rt_ExpFitEKF(c.x, c.s_init, c.params, c.w_bar, c.v_bar, c.Ps_init, c.Q_w, c.R_v, c.beta, c.gamma, ...
                 c.inv_monitor_len, c.order);
out.rt_expfit = o;

out.meta = struct('interpreter', version(), 'is_octave', exist('OCTAVE_VERSION', 'builtin') ~= 0);
if out.meta.is_octave
    save('-v7', out_file, '-struct', 'out');
else
    save(out_file, '-struct', 'out', '-v7');
end
fprintf('ref_fixtures: wrote %s\n', out_file);
end
