function r = randn(varargin)
% Deterministic stand-in for randn while oracle/ref_fixtures.m runs SIalpha_Controlled.m: returns the next
% value of the stream tests/cases.py drew (the reference calls randn with no argument, one scalar at a time,
% in the order s, i, alpha -- Tools/SIalpha_Controlled.m:25-27).  Test infrastructure; on the path only
% while ref_fixtures runs.
global EPI_REF_RANDN_STREAM EPI_REF_RANDN_POS
if nargin ~= 0
    error('ref_shims/randn: only the scalar form is replayed');
end
EPI_REF_RANDN_POS = EPI_REF_RANDN_POS + 1;
r = EPI_REF_RANDN_STREAM(EPI_REF_RANDN_POS);
end
