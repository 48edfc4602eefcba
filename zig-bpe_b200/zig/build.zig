// build.zig for the zig-bpe executable on top of libbpe_b200.so (replaces the raylib wiring of the
// reference's build.zig:7-27). UNVERIFIED: no Zig toolchain in this repository's build image.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC -shared \
//        -o zig-out/lib/libbpe_b200.so <b200-bpe>/zig-bpe_b200/csrc/bpe_engine.cu -ldl
const std = @import("std");

pub fn build(b: *std.Build) void {
    const target = b.standardTargetOptions(.{});
    const optimize = b.standardOptimizeOption(.{});
    const bpe_lib_dir = b.option([]const u8, "bpe-lib-dir", "directory holding libbpe_b200.so") orelse "zig-out/lib";
    const cuda_lib_dir = b.option([]const u8, "cuda-lib-dir", "CUDA runtime library directory") orelse "/usr/local/cuda/lib64";

    const exe = b.addExecutable(.{
        .name = "zig-bpe",
        .root_source_file = b.path("src/main.zig"),
        .target = target,
        .optimize = optimize,
    });
    exe.addLibraryPath(.{ .cwd_relative = bpe_lib_dir });
    exe.addLibraryPath(.{ .cwd_relative = cuda_lib_dir });
    exe.addRPath(.{ .cwd_relative = bpe_lib_dir });
    exe.linkSystemLibrary("bpe_b200");
    exe.linkSystemLibrary("cudart");
    exe.linkLibC();
    b.installArtifact(exe);

    const run_cmd = b.addRunArtifact(exe);
    run_cmd.step.dependOn(b.getInstallStep());
    const run_step = b.step("run", "Run the app");
    run_step.dependOn(&run_cmd.step);

    const tests = b.addTest(.{ .root_source_file = b.path("src/basic_tokenizer.zig"), .target = target, .optimize = optimize });
    tests.addLibraryPath(.{ .cwd_relative = bpe_lib_dir });
    tests.addRPath(.{ .cwd_relative = bpe_lib_dir });
    tests.linkSystemLibrary("bpe_b200");
    tests.linkLibC();
    const test_step = b.step("test", "Run the reference's in-file tests against the CUDA engine");
    test_step.dependOn(&b.addRunArtifact(tests).step);
}
