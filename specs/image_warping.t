-- 2-D as-rigid-as-possible image warp (opt_b200 test spec; same energy and parameter positions
-- as the image_warping example of niessner/Opt, written independently).
--   params: 0 Offset (unknown, 2ch)  1 Angle (unknown, 1ch)  2 UrShape  3 Constraints  4 Mask
--           5 w_fitSqrt (float)  6 w_regSqrt (float)
local W = Dim("W", 0)
local H = Dim("H", 1)
local dom = { W, H }
local Offset      = Unknown("Offset", opt_float2, dom, 0)
local Angle       = Unknown("Angle", opt_float, dom, 1)
local UrShape     = Array("UrShape", opt_float2, dom, 2)
local Constraints = Array("Constraints", opt_float2, dom, 3)
local Mask        = Array("Mask", opt_float, dom, 4)
local w_fitSqrt   = Param("w_fitSqrt", float, 5)
local w_regSqrt   = Param("w_regSqrt", float, 6)

UsePreconditioner(true)
-- masked-out pixels are not unknowns
Exclude(Not(eq(Mask(0, 0), 0)))

-- rigidity of every pixel against its 4-neighbourhood
local here = Offset(0, 0)
local rest = UrShape(0, 0)
for dx, dy in Stencil { { 1, 0 }, { -1, 0 }, { 0, 1 }, { 0, -1 } } do
    local rotated = Rotate2D(Angle(0, 0), rest - UrShape(dx, dy))
    local pairValid = InBounds(dx, dy) * eq(Mask(dx, dy), 0) * eq(Mask(0, 0), 0)
    Energy(Select(pairValid, w_regSqrt * ((here - Offset(dx, dy)) - rotated), 0))
end

-- soft positional constraints (negative coordinates mean "no constraint")
local pinned = All(greatereq(Constraints(0, 0), 0))
Energy(w_fitSqrt * Select(pinned, here - Constraints(0, 0), 0.0))
