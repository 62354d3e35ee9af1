-- As-rigid-as-possible warp of a 2-D image: every pixel carries a position and a rotation angle (test specification of
-- opt_b200, stated independently; the parameter table matches the image_warping example of niessner/Opt).
--   slot 0: pixel positions, 2 channels, UNKNOWN      slot 1: pixel rotation angle, UNKNOWN      slot 2: rest positions
--   slot 3: target positions (a negative coordinate = unconstrained pixel)      slot 4: mask (non-zero = pixel is not part of the problem)
--   slot 5 / 6: square roots of the constraint and rigidity weights
local pixels = { Dim("W", 0), Dim("H", 1) }
local function field(label, kind, slot) return Array(label, kind, pixels, slot) end

local where_  = Unknown("Offset", opt_float2, pixels, 0)
local turn    = Unknown("Angle", opt_float, pixels, 1)
local rest    = field("UrShape", opt_float2, 2)
local goal    = field("Constraints", opt_float2, 3)
local outside = field("Mask", opt_float, 4)
local sqrtPin, sqrtRigid = Param("w_fitSqrt", float, 5), Param("w_regSqrt", float, 6)
UsePreconditioner(true)

local function inside(dx, dy) return eq(outside(dx, dy), 0) end
Exclude(Not(inside(0, 0)))

-- a pixel and each of its four neighbours should move like a rigid body turned by the pixel's angle
local cross = { { 1, 0 }, { -1, 0 }, { 0, 1 }, { 0, -1 } }
for _, d in ipairs(cross) do
    local dx, dy = d[1], d[2]
    local restEdge = rest(0, 0) - rest(dx, dy)
    local nowEdge  = where_(0, 0) - where_(dx, dy)
    local usable   = InBounds(dx, dy) * inside(dx, dy) * inside(0, 0)
    Energy(Select(usable, sqrtRigid * (nowEdge - Rotate2D(turn(0, 0), restEdge)), 0))
end

-- pull constrained pixels to their targets
local hasGoal = All(greatereq(goal(0, 0), 0))
Energy(sqrtPin * Select(hasGoal, where_(0, 0) - goal(0, 0), 0.0))
