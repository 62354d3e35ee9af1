-- Two-parameter curve fit  y = a cos(b x) + b sin(a x)  stated purely over a graph: every edge joins one data sample with
-- THE parameter pair (test specification of opt_b200; parameter table of tests/minimal_graph_only in niessner/Opt).
--   slot 0: (a, b), UNKNOWN, index space U     slot 1: samples (x, y), index space N
--   slot 2: edge count     slot 3: sample index per edge     slot 4: parameter index per edge
local samples, unknowns = { Dim("N", 0) }, { Dim("U", 1) }
local coeff  = Unknown("funcParams", opt_float2, unknowns, 0)
local points = Array("data", opt_float2, samples, 1)
local pairs_ = Graph("G", 2, "d", samples, 3, "p", unknowns, 4)
UsePreconditioner(true)

local pt, ab = points(pairs_.d), coeff(pairs_.p)
local function model(x, a, b) return a * cos(b * x) + b * sin(a * x) end
Energy(pt(1) - model(pt(0), ab(0), ab(1)))
