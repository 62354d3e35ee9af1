-- Depth refinement from shading (test specification of opt_b200; the parameter table matches the shape_from_shading example of
-- niessner/Opt).  The rendered shading of the refined depth (second-order spherical harmonics of the surface normal) must
-- reproduce the intensity gradients of the colour image, while the depth stays near the sensor depth and locally smooth.
--   slots 0..2: weights of the depth, smoothness and shading terms (their square roots are used)
--   slots 3..6: camera intrinsics f_x, f_y, u_x, u_y          slots 7..15: nine lighting coefficients
--   slot 16: refined depth, UNKNOWN      slot 17: sensor depth (<= 0: no measurement)      slot 18: intensity
--   slots 19, 20: uint8 masks that switch the horizontal / vertical shading-gradient term off across image edges
local image = { Dim("W", 0), Dim("H", 1) }
local function scalar(label, slot) return Param(label, float, slot) end
local sqrtDepth, sqrtSmooth, sqrtShade = sqrt(scalar("w_p", 0)), sqrt(scalar("w_s", 1)), sqrt(scalar("w_g", 2))
local fx, fy, cx, cy = scalar("f_x", 3), scalar("f_y", 4), scalar("u_x", 5), scalar("u_y", 6)
local sh = {}
for k = 1, 9 do sh[k] = scalar("L_" .. k .. "", 6 + k) end
local depth     = Unknown("X", opt_float, image, 16)
local sensor    = Array("D_i", opt_float, image, 17)
local intensity = Array("Im", opt_float, image, 18)
local keepAcrossRows = Array("edgeMaskR", uint8, image, 19)
local keepAcrossCols = Array("edgeMaskC", uint8, image, 20)

local u, v = Index(0), Index(1)
local discontinuity = 0.01
local function measured(dx, dy) return greater(sensor(dx, dy), 0) end
Exclude(Not(measured(0, 0)))

-- unnormalised normal from the depths at a pixel, its left and its upper neighbour; then normalised where it has a length
local function unitNormal(dx, dy)
    local here, left, up = depth(dx, dy), depth(dx - 1, dy), depth(dx, dy - 1)
    local a = up * (here - left) / fy
    local b = left * (here - up) / fx
    local c = (a * (cx - (dx + u)) / fx) + (b * (cy - (dy + v)) / fy) - (left * up / (fx * fy))
    local sq = a * a + b * b + c * c
    return Select(greater(sq, 0.0), 1.0 / sqrt(sq), 1.0) * Vector(a, b, c)
end
local function rendered(dx, dy)
    local n = unitNormal(dx, dy)
    local a, b, c = n[0], n[1], n[2]
    return sh[1] + sh[2] * b + sh[3] * c + sh[4] * a + sh[5] * a * b + sh[6] * b * c
         + sh[7] * (-a * a - b * b + 2 * c * c) + sh[8] * c * a + sh[9] * (a * a - b * b)
end
local function observed(dx, dy) return intensity(dx, dy) * 0.5 + 0.25 * (intensity(dx - 1, dy) + intensity(dx, dy - 1)) end
local interior = InBoundsExpanded(0, 0, 1)
local function mismatch(dx, dy)
    local usable = measured(dx - 1, dy) * measured(dx, dy) * measured(dx, dy - 1)
    return Select(InBoundsExpanded(0, 0, 1) * usable, rendered(dx, dy) - observed(dx, dy), 0)
end
local shadingError = ComputedArray("B_I", image, mismatch(0, 0))

-- 1. depth prior
Energy(Select(measured(0, 0), sqrtDepth * (depth(0, 0) - sensor(0, 0)), 0))
-- 2. shading-gradient consistency, horizontal then vertical
Energy(Select(interior, sqrtShade * ((shadingError(0, 0) - shadingError(1, 0)) * keepAcrossRows(0, 0)), 0))
Energy(Select(interior, sqrtShade * ((shadingError(0, 0) - shadingError(0, 1)) * keepAcrossCols(0, 0)), 0))
-- 3. Laplacian of the back-projected surface where the depth has no discontinuity
local function backProject(dx, dy)
    local d = depth(dx, dy)
    return Vector((((dx + u) - cx) / fx) * d, (((dy + v) - cy) / fy) * d, d)
end
local function continuous(dx, dy) return less(abs(depth(0, 0) - depth(dx, dy)), discontinuity) end
local smoothHere = measured(0, 0) * measured(0, -1) * measured(0, 1) * measured(-1, 0) * measured(1, 0)
                 * continuous(0, -1) * continuous(0, 1) * continuous(-1, 0) * continuous(1, 0) * InBoundsExpanded(0, 0, 1)
local smoothFlag = ComputedArray("valid", image, smoothHere)
local laplace = 4.0 * backProject(0, 0) - (backProject(-1, 0) + backProject(0, -1) + backProject(1, 0) + backProject(0, 1))
Energy(Select(eq(smoothFlag(0, 0), 1), sqrtSmooth * laplace, 0))
