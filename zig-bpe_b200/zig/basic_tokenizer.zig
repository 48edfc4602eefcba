//! Drop-in replacement for dbtreasure/zig-bpe `src/basic_tokenizer.zig`: same public API
//! (`BasicTokenizer.{init,deinit,train,encode,decode,serializeMerges,deserializeMerges}`, `TrainError`,
//! `merges.put`, `generateInitialTokens`, `timeStats`), same allocator discipline, same file format and
//! stderr text. The hot loops (`expandVocabulary`, the encode pass loop, `decode`) call libbpe_b200.so
//! through the C ABI of include/bpe_b200.h.
//!
//! NOT COMPILED IN THIS REPOSITORY'S CI: no Zig 0.13 toolchain exists in the build image. The C ABI it
//! binds is exercised by the Python and C++ mirrors (tests/, zig-bpe_b200/host/).
const std = @import("std");

const TimeStats = @import("utils/time_statistics.zig").TimeStats;
const printTimeStats = @import("utils/time_statistics.zig").printTimeStats;

pub const TrainError = error{
    InvalidVocabSize,
    InvalidUtf8,
    OutOfMemory,
};

// ---- C ABI (include/bpe_b200.h) -------------------------------------------------------------
const bpe_ctx = opaque {};
const bpe_merge_t = extern struct { first: u16, second: u16, new_token: u16 };
const bpe_stats_t = extern struct {
    sort_pairs_ms: f64,
    replace_pair_ms: f64,
    generate_pairs_ms: f64,
    just_count_pairs_ms: f64,
    sort_pairs_calls: u64,
    replace_pair_calls: u64,
    generate_pairs_calls: u64,
    just_count_pairs_calls: u64,
    total_ms: f64,
    device_ms: f64,
    scanned_slots: u64,
    kernel_launches: u64,
    tie_steps: u64,
    tie_slow_steps: u64,
    compactions: u64,
    kernel_ms: [12]f64,
    kernel_calls: [12]u64,
    aeqb_steps: u64,
};
extern fn bpe_ctx_create(out: *?*bpe_ctx, device: c_int) c_int;
extern fn bpe_ctx_destroy(ctx: ?*bpe_ctx) void;
extern fn bpe_ctx_set_option(ctx: ?*bpe_ctx, name: [*:0]const u8, value: c_long) c_int;
extern fn bpe_train(ctx: ?*bpe_ctx, text: [*]const u8, n: usize, vocab_size: u16, out_merges: [*]bpe_merge_t, out_counts: ?[*]u64, out_n: *usize, stats: ?*bpe_stats_t) c_int;
extern fn bpe_encode(ctx: ?*bpe_ctx, text: [*]const u8, n: usize, merges: [*]const bpe_merge_t, m: usize, out: [*]u16, out_n: *usize, stats: ?*bpe_stats_t) c_int;
extern fn bpe_decode_size(ctx: ?*bpe_ctx, toks: [*]const u16, n: usize, merges: [*]const bpe_merge_t, m: usize, out_n: *usize) c_int;
extern fn bpe_decode(ctx: ?*bpe_ctx, toks: [*]const u16, n: usize, merges: [*]const bpe_merge_t, m: usize, out: [*]u8, cap: usize, out_n: *usize, stats: ?*bpe_stats_t) c_int;

const BPE_OK = 0;
const BPE_ERR_INVALID_VOCAB = 1;
const BPE_ERR_OOM = 2;
const BPE_ERR_INVALID_TOKEN = 3;

const Merge = struct {
    pair: CharPair,
    new_token: u16,
};

const Merges = struct {
    merges: std.ArrayList(Merge),
    allocator: std.mem.Allocator,

    pub fn init(allocator: std.mem.Allocator) @This() {
        return .{ .merges = std.ArrayList(Merge).init(allocator), .allocator = allocator };
    }
    pub fn deinit(self: *Merges) void {
        self.merges.deinit();
    }
    pub fn put(self: *Merges, pair: CharPair, new_token: u16) !void {
        try self.merges.append(.{ .pair = pair, .new_token = new_token });
    }
};

const CharPair = struct {
    first: u16,
    second: u16,
};

const vocabStart: u16 = 256;

pub const BasicTokenizer = struct {
    allocator: std.mem.Allocator,
    timeStats: *TimeStats,
    merges: Merges,
    ctx: ?*bpe_ctx,

    pub fn init(allocator: std.mem.Allocator) !@This() {
        const timeStats = try TimeStats.init(allocator);
        errdefer timeStats.deinit();
        var ctx: ?*bpe_ctx = null;
        // CUDA failure has no reference equivalent; OutOfMemory is the closest member of the error set
        if (bpe_ctx_create(&ctx, 0) != BPE_OK) return error.OutOfMemory;
        _ = bpe_ctx_set_option(ctx, "time_phases", 1);
        return .{ .allocator = allocator, .timeStats = timeStats, .merges = Merges.init(allocator), .ctx = ctx };
    }

    pub fn deinit(self: *@This()) void {
        bpe_ctx_destroy(self.ctx);
        self.timeStats.deinit();
        self.merges.deinit();
    }

    /// the merge list in the ABI's 6-byte layout (caller frees)
    fn abiMerges(self: *@This()) ![]bpe_merge_t {
        const out = try self.allocator.alloc(bpe_merge_t, self.merges.merges.items.len);
        for (self.merges.merges.items, 0..) |m, i| out[i] = .{ .first = m.pair.first, .second = m.pair.second, .new_token = m.new_token };
        return out;
    }

    pub fn encode(self: *@This(), text: []const u8) !std.ArrayList(u16) {
        var tokens = std.ArrayList(u16).init(self.allocator);
        errdefer tokens.deinit();
        try tokens.resize(text.len); // upper bound: one id per byte
        const ms = try self.abiMerges();
        defer self.allocator.free(ms);
        var n: usize = 0;
        const rc = bpe_encode(self.ctx, text.ptr, text.len, ms.ptr, ms.len, tokens.items.ptr, &n, null);
        if (rc != BPE_OK) return error.OutOfMemory;
        tokens.shrinkRetainingCapacity(n);
        return tokens;
    }

    pub fn decode(self: *@This(), tokens: std.ArrayList(u16)) ![]u8 {
        const ms = try self.abiMerges();
        defer self.allocator.free(ms);
        var need: usize = 0;
        var rc = bpe_decode_size(self.ctx, tokens.items.ptr, tokens.items.len, ms.ptr, ms.len, &need);
        if (rc == BPE_ERR_INVALID_TOKEN) return error.InvalidToken;
        if (rc != BPE_OK) return error.OutOfMemory;
        const out = try self.allocator.alloc(u8, need);
        errdefer self.allocator.free(out);
        var n: usize = 0;
        rc = bpe_decode(self.ctx, tokens.items.ptr, tokens.items.len, ms.ptr, ms.len, out.ptr, need, &n, null);
        if (rc == BPE_ERR_INVALID_TOKEN) return error.InvalidToken;
        if (rc != BPE_OK) return error.OutOfMemory;
        return out;
    }

    pub fn train(self: *@This(), text: []const u8, vocabSize: u16, verbose: bool) TrainError!void {
        const start = std.time.milliTimestamp();
        defer {
            const end = std.time.milliTimestamp();
            printTimeStats(self.timeStats, end - start);
        }
        if (vocabSize < 256) return TrainError.InvalidVocabSize;
        const cap: usize = vocabSize - vocabStart;
        const out = try self.allocator.alloc(bpe_merge_t, @max(cap, 1));
        defer self.allocator.free(out);
        const counts = try self.allocator.alloc(u64, @max(cap, 1));
        defer self.allocator.free(counts);
        var n: usize = 0;
        var st: bpe_stats_t = undefined;
        const rc = bpe_train(self.ctx, text.ptr, text.len, vocabSize, out.ptr, counts.ptr, &n, &st);
        if (rc == BPE_ERR_INVALID_VOCAB) return TrainError.InvalidVocabSize;
        if (rc != BPE_OK) return TrainError.OutOfMemory;
        // (explicit @as: compound assignment gives the cast builtins no result type)
        self.timeStats.sort_pairs_time += @as(i64, @intFromFloat(st.sort_pairs_ms));
        self.timeStats.sort_pairs_calls += @as(usize, @intCast(st.sort_pairs_calls));
        self.timeStats.replace_pair_time += @as(i64, @intFromFloat(st.replace_pair_ms));
        self.timeStats.replace_pair_calls += @as(usize, @intCast(st.replace_pair_calls));
        self.timeStats.generate_pairs_time += @as(i64, @intFromFloat(st.generate_pairs_ms));
        self.timeStats.generate_pairs_calls += @as(usize, @intCast(st.generate_pairs_calls));
        self.timeStats.just_count_pairs_time += @as(i64, @intFromFloat(st.just_count_pairs_ms));
        self.timeStats.just_count_pairs_calls += @as(usize, @intCast(st.just_count_pairs_calls));
        for (out[0..n], 0..) |m, i| {
            if (verbose) {
                std.debug.print("merge {d}/{d}: ({d},{d}) -> {d} had {d} occurrences\n", .{ i + 1, vocabSize - vocabStart, m.first, m.second, m.new_token, counts[i] });
            }
            try self.merges.put(.{ .first = m.first, .second = m.second }, m.new_token);
        }
        if (n < cap) std.debug.print("No more pairs to merge. Stopping early.\n", .{});
    }

    pub fn generateInitialTokens(self: *BasicTokenizer, text: []const u8) TrainError!std.ArrayList(u16) {
        var tokens = std.ArrayList(u16).init(self.allocator);
        errdefer tokens.deinit();
        for (text) |byte| try tokens.append(@as(u16, byte));
        return tokens;
    }

    pub fn serializeMerges(self: *@This(), file_path: []const u8) !void {
        const file = try std.fs.cwd().createFile(file_path, .{});
        defer file.close();
        var writer = file.writer();
        for (self.merges.merges.items) |entry| {
            try writer.print("{d},{d},{d}\n", .{ entry.pair.first, entry.pair.second, entry.new_token });
        }
    }

    pub fn deserializeMerges(self: *@This(), file_path: []const u8) !void {
        const file = try std.fs.cwd().openFile(file_path, .{});
        defer file.close();
        var buf_reader = std.io.bufferedReader(file.reader());
        var in_stream = buf_reader.reader();
        var buf: [100]u8 = undefined;
        while (try in_stream.readUntilDelimiterOrEof(&buf, '\n')) |line| {
            var it = std.mem.split(u8, line, ",");
            const first = try std.fmt.parseInt(u16, it.next() orelse return error.InvalidFormat, 10);
            const second = try std.fmt.parseInt(u16, it.next() orelse return error.InvalidFormat, 10);
            const new_token = try std.fmt.parseInt(u16, it.next() orelse return error.InvalidFormat, 10);
            try self.merges.put(.{ .first = first, .second = second }, new_token);
        }
    }
};
