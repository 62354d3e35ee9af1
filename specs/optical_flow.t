-- variational optical flow with a bilinearly sampled target image (opt_b200 test spec; same energy and parameter
-- positions as the optical_flow example of niessner/Opt).  Served by the general front end.
--   params: 0 w_fit  1 w_reg  2 X (unknown flow, 2ch)  3 I  4 I_hat  5 I_hat_dx  6 I_hat_dy
local W = Dim("W", 0)
local H = Dim("H", 1)
local w_fitSqrt = Param("w_fit", float, 0)
local w_regSqrt = Param("w_reg", float, 1)
local X        = Unknown("X", opt_float2, { W, H }, 2)
local I        = Array("I", opt_float, { W, H }, 3)
local I_hat_im = Array("I_hat", opt_float, { W, H }, 4)
local I_hat_dx = Array("I_hat_dx", opt_float, { W, H }, 5)
local I_hat_dy = Array("I_hat_dy", opt_float, { W, H }, 6)
local warped = SampledImage(I_hat_im, I_hat_dx, I_hat_dy)
local col, row = Index(0), Index(1)
UsePreconditioner(false)

-- brightness constancy at the displaced position
Energy(w_fitSqrt * (I(0, 0) - warped(col + X(0, 0, 0), row + X(0, 0, 1))))
-- first-order smoothness of the flow
for dx, dy in Stencil { { 1, 0 }, { -1, 0 }, { 0, 1 }, { 0, -1 } } do
    Energy(Select(InBounds(dx, dy), w_regSqrt * (X(0, 0) - X(dx, dy)), 0))
end
