# Julia-side test file mirroring the reference's test/cuda.jl (GPU result ≈ CPU result) for the libdpr-backed
# extension.  Not runnable in the build image (no Julia); the same comparisons run in tests/test_gpu_parity.py.
@testitem "B200 forward" begin
    using Adapt, CUDA
    CUDA.allowscalar(false)
    include("data.jl")
    include("util.jl")
    ok = CUDA.functional()
    for args in (
        (D.grid_size_3d, D.more_points, D.rotations_static, D.translations_3d_static, D.backgrounds, D.weights, D.more_point_weights),
        (D.grid_size_3d, D.more_points, D.rotations_static, D.translations_3d_static),
        (D.grid_size_2d, D.more_points, D.projections_static, D.translations_2d_static, D.backgrounds, D.weights, D.more_point_weights),
    )
        @test cuda_cpu_agree(raster, args...) skip = !ok
    end
end

@testitem "B200 backward" begin
    using Adapt, CUDA
    CUDA.allowscalar(false)
    include("data.jl")
    include("util.jl")
    ok = CUDA.functional()
    ds_dout_3d = randn(D.grid_size_3d..., D.batch_size)
    ds_dout_2d = randn(D.grid_size_2d..., D.batch_size)
    for args in (
        (ds_dout_3d, D.more_points, D.rotations_static, D.translations_3d_static, D.backgrounds, D.weights, D.more_point_weights),
        (ds_dout_3d, D.more_points, D.rotations_static, D.translations_3d_static),
        (ds_dout_2d, D.more_points, D.projections_static, D.translations_2d_static, D.backgrounds, D.weights, D.more_point_weights),
    )
        @test cuda_cpu_agree(raster_pullback!, args...) skip = !ok
    end
end
