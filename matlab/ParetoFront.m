function [on_front, I_opt] = ParetoFront(J0, J1)
% The strict-dominance filter and knee point of Tools/TrainPredictPrescribeNPI.m:624-633.
[on_front, I_opt] = epi_mex('pareto', J0(:)', J1(:)');
end
