-- Seamless cloning: inside the blend region the result must have the gradients of the inserted image (test specification of
-- opt_b200; the parameter table matches the poisson_image_editing example of niessner/Opt).
--   slot 0: result, 4 channels, UNKNOWN      slot 1: inserted image, 4 channels      slot 2: mask (0 = inside the blend region)
local canvas = { Dim("W", 0), Dim("H", 1) }
local result  = Unknown("X", opt_float4, canvas, 0)
local source  = Array("T", opt_float4, canvas, 1)
local regionMask = Array("M", opt_float, canvas, 2)
UsePreconditioner(false)
Exclude(Not(eq(regionMask(0, 0), 0)))

local function difference(img, dx, dy) return img(0, 0) - img(dx, dy) end
for _, d in ipairs({ { 1, 0 }, { -1, 0 }, { 0, 1 }, { 0, -1 } }) do
    local mismatch = difference(result, d[1], d[2]) - difference(source, d[1], d[2])
    Energy(Select(InBounds(d[1], d[2]), mismatch, 0))
end
