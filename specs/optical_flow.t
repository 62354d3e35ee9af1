-- Dense optical flow between two frames (test specification of opt_b200; the parameter table matches the optical_flow
-- example of niessner/Opt, so the example's host code binds to it unchanged).  No built-in kernel set: general front end.
--   slot 0: weight of the data term (already a square root)      slot 1: weight of the smoothness term (square root)
--   slot 2: flow field, two channels, UNKNOWN                    slot 3: first frame
--   slot 4..6: second frame and its x / y derivative images (sampled bilinearly at the displaced position)
local grid = { Dim("W", 0), Dim("H", 1) }
local function frame(label, slot) return Array(label, opt_float, grid, slot) end

local lambdaData   = Param("w_fit", float, 0)
local lambdaSmooth = Param("w_reg", float, 1)
local flow   = Unknown("X", opt_float2, grid, 2)
local first  = frame("I", 3)
local second = SampledImage(frame("I_hat", 4), frame("I_hat_dx", 5), frame("I_hat_dy", 6))
UsePreconditioner(false)

-- brightness constancy: the first frame at a pixel equals the second frame at the pixel moved by the flow
local px, py = Index(0), Index(1)
local u, v = flow(0, 0, 0), flow(0, 0, 1)
Energy(lambdaData * (first(0, 0) - second(px + u, py + v)))

-- membrane smoothness towards the four axis neighbours that exist
local shifts = { { 1, 0 }, { -1, 0 }, { 0, 1 }, { 0, -1 } }
for _, s in ipairs(shifts) do
    local jump = flow(0, 0) - flow(s[1], s[2])
    Energy(Select(InBounds(s[1], s[2]), lambdaSmooth * jump, 0))
end
