-- shape-from-shading depth refinement (opt_b200 test spec; same energy and parameter positions as the
-- shape_from_shading example of niessner/Opt).
--   params: 0 w_p  1 w_s  2 w_g  3 f_x  4 f_y  5 u_x  6 u_y  7..15 L_1..L_9 (all float)
--           16 X (unknown depth)  17 D_i  18 Im  19 edgeMaskR (uint8)  20 edgeMaskC (uint8)
local W = Dim("W", 0)
local H = Dim("H", 1)
local dom = { W, H }
local sqrt_w_p = sqrt(Param("w_p", float, 0))
local sqrt_w_s = sqrt(Param("w_s", float, 1))
local sqrt_w_g = sqrt(Param("w_g", float, 2))
local f_x = Param("f_x", float, 3)
local f_y = Param("f_y", float, 4)
local u_x = Param("u_x", float, 5)
local u_y = Param("u_y", float, 6)
local light = {}
for k = 1, 9 do
    light[k] = Param("L_" .. k .. "", float, 6 + k)
end
local X         = Unknown("X", opt_float, dom, 16)
local D_i       = Array("D_i", opt_float, dom, 17)
local Im        = Array("Im", opt_float, dom, 18)
local edgeMaskR = Array("edgeMaskR", uint8, dom, 19)
local edgeMaskC = Array("edgeMaskC", uint8, dom, 20)

local col, row = Index(0), Index(1)
local JUMP = 0.01

local function hasDepth(dx, dy) return greater(D_i(dx, dy), 0) end

-- back-projected point of pixel (dx,dy)
local function point(dx, dy)
    local depth = X(dx, dy)
    return Vector((((dx + col) - u_x) / f_x) * depth, (((dy + row) - u_y) / f_y) * depth, depth)
end

-- unit normal from depth differences
local function normal(dx, dy)
    local i = dx + col
    local j = dy + row
    local nx = X(dx, dy - 1) * (X(dx, dy) - X(dx - 1, dy)) / f_y
    local ny = X(dx - 1, dy) * (X(dx, dy) - X(dx, dy - 1)) / f_x
    local nz = (nx * (u_x - i) / f_x) + (ny * (u_y - j) / f_y) - (X(dx - 1, dy) * X(dx, dy - 1) / (f_x * f_y))
    local len2 = nx * nx + ny * ny + nz * nz
    local scale = Select(greater(len2, 0.0), 1.0 / sqrt(len2), 1.0)
    return scale * Vector(nx, ny, nz)
end

-- second-order spherical-harmonics shading
local function shading(dx, dy)
    local n = normal(dx, dy)
    local nx, ny, nz = n[0], n[1], n[2]
    return light[1] + light[2] * ny + light[3] * nz + light[4] * nx
         + light[5] * nx * ny + light[6] * ny * nz + light[7] * (-nx * nx - ny * ny + 2 * nz * nz)
         + light[8] * nz * nx + light[9] * (nx * nx - ny * ny)
end

local function intensity(dx, dy)
    return Im(dx, dy) * 0.5 + 0.25 * (Im(dx - 1, dy) + Im(dx, dy - 1))
end

local function shadingError(dx, dy)
    local ok = hasDepth(dx - 1, dy) * hasDepth(dx, dy) * hasDepth(dx, dy - 1)
    return Select(InBoundsExpanded(0, 0, 1) * ok, shading(dx, dy) - intensity(dx, dy), 0)
end
local B_I = ComputedArray("B_I", dom, shadingError(0, 0))

Exclude(Not(hasDepth(0, 0)))

-- stay close to the measured depth
Energy(Select(hasDepth(0, 0), sqrt_w_p * (X(0, 0) - D_i(0, 0)), 0))

-- shading gradients must match intensity gradients (except across masked edges)
local inner = InBoundsExpanded(0, 0, 1)
Energy(Select(inner, sqrt_w_g * ((B_I(0, 0) - B_I(1, 0)) * edgeMaskR(0, 0)), 0))
Energy(Select(inner, sqrt_w_g * ((B_I(0, 0) - B_I(0, 1)) * edgeMaskC(0, 0)), 0))

-- Laplacian smoothness of the back-projected surface where the depth is continuous
local function smooth(dx, dy) return less(abs(X(0, 0) - X(dx, dy)), JUMP) end
local flat = hasDepth(0, 0) * hasDepth(0, -1) * hasDepth(0, 1) * hasDepth(-1, 0) * hasDepth(1, 0)
           * smooth(0, -1) * smooth(0, 1) * smooth(-1, 0) * smooth(1, 0) * InBoundsExpanded(0, 0, 1)
local flatArray = ComputedArray("valid", dom, flat)
local lap = 4.0 * point(0, 0) - (point(-1, 0) + point(0, -1) + point(1, 0) + point(0, 1))
Energy(Select(eq(flatArray(0, 0), 1), sqrt_w_s * lap, 0))
