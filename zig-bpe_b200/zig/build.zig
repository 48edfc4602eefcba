// build.zig for the zig-bpe executable on top of libbpe_b200.so (replaces the raylib wiring of the
// reference's build.zig:7-27). UNVERIFIED: no Zig toolchain in this repository's build image.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC -shared \
//        -o zig-out/lib/libbpe_b200.so <b200-bpe>/zig-bpe_b200/csrc/bpe_engine.cu -ldl
const std = @import("std");

const EngineDirs = struct { bpe: []const u8, cuda: []const u8 };

/// link a compile step (the executable or the test binary) against the CUDA engine
fn linkEngine(step: *std.Build.Step.Compile, dirs: EngineDirs) void {
    for ([_][]const u8{ dirs.bpe, dirs.cuda }) |dir| step.addLibraryPath(.{ .cwd_relative = dir });
    step.addRPath(.{ .cwd_relative = dirs.bpe });
    step.linkSystemLibrary("bpe_b200");
    step.linkSystemLibrary("cudart");
    step.linkLibC();
}

pub fn build(b: *std.Build) void {
    const dirs = EngineDirs{
        .bpe = b.option([]const u8, "bpe-lib-dir", "directory holding libbpe_b200.so") orelse "zig-out/lib",
        .cuda = b.option([]const u8, "cuda-lib-dir", "CUDA runtime library directory") orelse "/usr/local/cuda/lib64",
    };
    const opts = .{ .target = b.standardTargetOptions(.{}), .optimize = b.standardOptimizeOption(.{}) };

    // `zig build run`: the main.zig workload (train 300, serialize, encode, decode)
    const app = b.addExecutable(.{ .name = "zig-bpe", .root_source_file = b.path("src/main.zig"), .target = opts.target, .optimize = opts.optimize });
    linkEngine(app, dirs);
    b.installArtifact(app);
    const launch = b.addRunArtifact(app);
    launch.step.dependOn(b.getInstallStep());
    b.step("run", "Train, serialize, encode and decode through the CUDA engine").dependOn(&launch.step);

    // `zig build test`: the reference's in-file tests against the CUDA engine
    const unit = b.addTest(.{ .root_source_file = b.path("src/basic_tokenizer.zig"), .target = opts.target, .optimize = opts.optimize });
    linkEngine(unit, dirs);
    b.step("test", "Run the reference's in-file tests against the CUDA engine").dependOn(&b.addRunArtifact(unit).step);
}
