# Julia-side checks for the libdpr-backed extension: device results must agree with the package's CPU methods.
# (Cannot run in the build image - no Julia; tests/test_gpu_parity.py makes the same comparisons through ctypes.)
#
# Usage from the package's test environment:  include("test_b200.jl"); B200Checks.run()
module B200Checks

using Test, CUDA, Adapt, StaticArrays, Rotations, DiffPointRasterisation

to_device(x) = adapt(CuArray, x)
host(x) = adapt(Array, x)

"relative L2 distance of two arrays / numbers"
reldist(a, b) = (d = sqrt(sum(abs2, host(a) .- b)); n = sqrt(sum(abs2, b)); n > 0 ? d / n : d)

function fixtures(::Type{T}, n_out; n_points=50_000, batch=7) where {T}
    pts = [T(0.4) * @SVector(randn(T, 3)) for _ in 1:n_points]
    rots = [SMatrix{n_out,3,T}(rand(RotMatrix3{T})[1:n_out, :]) for _ in 1:batch]
    trs = [T(0.1) * @SVector(randn(T, n_out)) for _ in 1:batch]
    bg = collect(T, 1:batch)
    ow = T(10) .* rand(T, batch)
    pw = (w = rand(T, n_points); w ./ sum(w))
    return (; pts, rots, trs, bg, ow, pw)
end

function run(; tol32=1f-4, tol64=1e-9)
    CUDA.functional() || (@info "no CUDA device: skipping"; return)
    CUDA.allowscalar(false)
    @testset "libdpr vs CPU methods" begin
        for T in (Float32, Float64), (n_out, grid) in ((2, (32, 24)), (3, (16, 12, 10))), with_optional in (true, false)
            f = fixtures(T, n_out)
            args = with_optional ? (f.pts, f.rots, f.trs, f.bg, f.ow, f.pw) : (f.pts, f.rots, f.trs)
            tol = T === Float32 ? tol32 : tol64
            out_cpu = raster(grid, args...)
            out_gpu = raster(grid, map(to_device, args)...)
            @test reldist(out_gpu, out_cpu) <= tol
            ds_dout = randn(T, grid..., length(f.rots))
            pb_cpu = raster_pullback!(ds_dout, args...)
            pb_gpu = raster_pullback!(to_device(ds_dout), map(to_device, args)...)
            for name in propertynames(pb_cpu)
                @test reldist(getproperty(pb_gpu, name), getproperty(pb_cpu, name)) <= tol
            end
        end
    end
end

end # module
