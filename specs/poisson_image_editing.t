-- Poisson image blending (opt_b200 test spec; same energy and parameter positions as the
-- poisson_image_editing example of niessner/Opt).  params: 0 X (unknown, 4ch)  1 T  2 M
local W = Dim("W", 0)
local H = Dim("H", 1)
local X = Unknown("X", opt_float4, { W, H }, 0)
local T = Array("T", opt_float4, { W, H }, 1)
local M = Array("M", opt_float, { W, H }, 2)
UsePreconditioner(false)
Exclude(Not(eq(M(0, 0), 0)))
-- gradients of the result must match gradients of the inserted image
for dx, dy in Stencil { { 1, 0 }, { -1, 0 }, { 0, 1 }, { 0, -1 } } do
    local gradX = X(0, 0) - X(dx, dy)
    local gradT = T(0, 0) - T(dx, dy)
    Energy(Select(InBounds(dx, dy), gradX - gradT, 0))
end
