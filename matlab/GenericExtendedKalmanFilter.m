function varargout = GenericExtendedKalmanFilter(u, x, handles, params, s_init, Ps_init, s_final, Ps_final, w_bar, v_bar, Q_w, R_v, beta, gamma, inv_monitor_len, order)
% Drop-in for Tools/GenericExtendedKalmanFilter.m:1.  The reference accepts arbitrary
% function handles; only the reference's own forward handle sets can run on the device.
% They are recognised by the name of handles.NlinStateUpdate; anything else falls through
% to the original MATLAB implementation (keep it on the path as GenericExtendedKalmanFilter_ref).
name = func2str(handles.NlinStateUpdate);
if(numel(s_init) == 3 && ~isempty(strfind(name, 'NlinStateUpdate')) && isempty(strfind(name, 'Flipped')))
    model = 0;
elseif(numel(s_init) == 6 && ~isempty(strfind(name, 'NlinStateUpdate')) && isempty(strfind(name, 'Flipped')))
    model = 2;
else
    [varargout{1:nargout}] = GenericExtendedKalmanFilter_ref(u, x, handles, params, s_init, Ps_init, s_final, Ps_final, w_bar, v_bar, Q_w, R_v, beta, gamma, inv_monitor_len, order);
    return;
end
[varargout{1:max(nargout, 1)}] = epi_mex('ekf_eks', model, u, x(:)', params, s_init(:), Ps_init, s_final(:), Ps_final, w_bar, v_bar, Q_w, R_v, beta, gamma, inv_monitor_len, order);
end
