-- smoothing toy problem (opt_b200 test spec; same energy as tests/minimal of niessner/Opt)
local W = Dim("W", 0)
local H = Dim("H", 1)
local X = Unknown("X", float, { W, H }, 0)
local A = Array("A", float, { W, H }, 1)
local w_fit = 0.2
Energy(w_fit * (X(0, 0) - A(0, 0)))
Energy(X(0, 0) - X(1, 0))
Energy(X(0, 0) - X(0, 1))
