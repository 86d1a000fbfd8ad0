function varargout = GenericExtendedKalmanFilter(u, x, handles, params, s_init, Ps_init, s_final, Ps_final, w_bar, v_bar, Q_w, R_v, beta, gamma, inv_monitor_len, order)
% Drop-in for Tools/GenericExtendedKalmanFilter.m:1.  The reference accepts ARBITRARY function handles; only
% the reference's own forward callback sets exist on the device.  A call takes the device path only when all
% eight callbacks are, by identity, the local functions of ONE of the reference's forward wrapper files
% (Tools/SIAlphaModelEKF.m, Tools/SIAlphaModelEKFOptControlled.m) sitting next to the original
% implementation (kept on the path as GenericExtendedKalmanFilter_ref.m) -- see epi_known_handle_set below.
% Everything else -- a user model built by copying a wrapper and editing its callbacks (whatever the
% callbacks are called), anonymous functions, the time-reversed "Flipped" sets, which reach this function
% with already reversed inputs -- runs the original MATLAB implementation, unchanged.
% A wrapper file edited IN PLACE is indistinguishable by identity: delete its shim from matlab/ and set the
% environment variable EPI_B200_GENERIC_REF=1 (forces the pass-through) if you modify the reference's models.
model = epi_known_handle_set(handles, numel(s_init));
if(model < 0)
    [varargout{1:nargout}] = GenericExtendedKalmanFilter_ref(u, x, handles, params, s_init, Ps_init, s_final, Ps_final, w_bar, v_bar, Q_w, R_v, beta, gamma, inv_monitor_len, order);
    return;
end
[varargout{1:max(nargout, 1)}] = epi_mex('ekf_eks', model, u, x(:)', params, s_init(:), Ps_init, s_final(:), Ps_final, w_bar, v_bar, Q_w, R_v, beta, gamma, inv_monitor_len, order);
end

function model = epi_known_handle_set(handles, m)
% EPI_MODEL id (include/epi_b200.h) of the callback set, or -1.  Identity, not names: functions(h).file must
% be the same file for all eight handles, that file must be <reference Tools>/SIAlphaModelEKF.m or
% <reference Tools>/SIAlphaModelEKFOptControlled.m, and every handle must be the local function of that file
% the wrapper itself installs (SIAlphaModelEKF.m:9-16).
model = -1;
if(~isempty(getenv('EPI_B200_GENERIC_REF'))), return; end
fields = {'StateHardMargins', 'ObsHardMargins', 'NlinStateUpdate', 'NlinObsUpdate', 'StateJacobians', 'ObsJacobian', 'StateHessianTerms', 'ObsHessianTerms'};
wrappers = {'SIAlphaModelEKF', 0, 3; 'SIAlphaModelEKFOptControlled', 2, 6};
ref_impl = which('GenericExtendedKalmanFilter_ref');
if(isempty(ref_impl) || ~isstruct(handles)), return; end
ref_dir = fileparts(ref_impl);
file0 = '';
row = 0;
for k = 1 : numel(fields)
    if(~isfield(handles, fields{k})), return; end
    h = handles.(fields{k});
    if(~isa(h, 'function_handle')), return; end
    info = functions(h);
    if(~isfield(info, 'file') || isempty(info.file) || ~isfield(info, 'function')), return; end
    [d, base] = fileparts(info.file);
    if(k == 1)
        file0 = info.file;
        row = find(strcmp(wrappers(:, 1), base));
        if(isempty(row) || ~strcmp(d, ref_dir)), return; end
    elseif(~strcmp(info.file, file0))
        return;
    end
    fname = info.function;                       % 'NlinStateUpdate' (MATLAB) or 'SIAlphaModelEKF>NlinStateUpdate'
    gt = find(fname == '>', 1, 'last');
    if(~isempty(gt)), fname = fname(gt + 1 : end); end
    if(~strcmp(fname, fields{k})), return; end
end
if(m ~= wrappers{row, 3}), return; end
model = wrappers{row, 2};
end
