//! TimeStats of the reference (src/utils/time_statistics.zig), kept on the host side: the buckets are
//! filled from the CUDA-event timings the C ABI returns in bpe_stats_t. UNVERIFIED (no Zig toolchain here).
const std = @import("std");
const Allocator = std.mem.Allocator;

pub const TimeStats = struct {
    allocator: Allocator,
    sort_pairs_time: i64 = 0,
    sort_pairs_calls: usize = 0,
    replace_pair_time: i64 = 0,
    replace_pair_calls: usize = 0,
    generate_pairs_time: i64 = 0,
    generate_pairs_calls: usize = 0,
    just_count_pairs_time: i64 = 0,
    just_count_pairs_calls: usize = 0,

    pub fn init(allocator: Allocator) !*TimeStats {
        const self = try allocator.create(TimeStats);
        self.* = .{ .allocator = allocator };
        return self;
    }

    pub fn deinit(self: *TimeStats) void {
        self.allocator.destroy(self);
    }
};

fn line(name: []const u8, t: i64, calls: usize) void {
    const tf = @as(f64, @floatFromInt(t));
    std.debug.print("{s}: {d:.3}s total, {d} calls, {d:.3}s avg\n", .{ name, tf / 1000.0, calls, tf / (@as(f64, @floatFromInt(calls)) * 1000.0) });
}

pub fn printTimeStats(stats: *const TimeStats, total_time: i64) void {
    std.debug.print("\nTime statistics:\n", .{});
    line("sortCodePointPairs", stats.sort_pairs_time, stats.sort_pairs_calls);
    line("replaceTopPairWithIndex", stats.replace_pair_time, stats.replace_pair_calls);
    line("generateCodePointPairs", stats.generate_pairs_time, stats.generate_pairs_calls);
    line("countPointPairs", stats.just_count_pairs_time, stats.just_count_pairs_calls);
    const other = total_time - stats.sort_pairs_time - stats.replace_pair_time - stats.generate_pairs_time - stats.just_count_pairs_time;
    std.debug.print("Other operations: {d:.3}s\n", .{@as(f64, @floatFromInt(other)) / 1000.0});
}
