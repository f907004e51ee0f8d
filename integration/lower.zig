//! lower.zig — the host-side shim that makes librtb.so a drop-in for the 8 render threads of
//! dariooddenino/zig-raytracing-weekend: lowers `world` (the BVHNode pointer graph and the Hittable / Material /
//! Texture tagged unions) into the POD arrays of include/rtb.h ONCE per scene, lowers the Camera, and replaces
//! startRender / stopRender / shouldStopRender (src/main.zig:314-348).
//!
//! NOT COMPILED here (no Zig toolchain in the build image; see rtb.zig).  It is written against the reference's HEAD
//! (field names as of src/objects.zig, src/material.zig, src/textures.zig, src/perlin.zig, src/bvh.zig, src/camera.zig)
//! and mirrors, function for function, zig-raytracing-weekend_b200/host/rtw_host.cpp — the C++ twin that IS compiled and
//! whose output is checked bit for bit against the oracle in tests/test_host_and_abi.py and tests/test_cornell.py.
const std = @import("std");
const rtb = @import("rtb.zig");
const objects = @import("objects.zig");
const materials = @import("material.zig");
const textures = @import("textures.zig");
const bvh = @import("bvh.zig");
const cameras = @import("camera.zig");
const zstbi = @import("zstbi");

const Hittable = objects.Hittable;
const Material = materials.Material;
const Texture = textures.Texture;
const BVHNode = bvh.BVHNode;
const Camera = cameras.Camera;
const Vec3 = @import("vec3.zig").Vec3;

fn v3(v: Vec3) [3]f32 {
    return .{ v[0], v[1], v[2] };
}

pub const LowerError = error{ Unsupported, OutOfMemory };

/// The POD arrays handed to rtb_scene_create.  They may be freed as soon as that call returns (the library copies).
pub const Lowered = struct {
    allocator: std.mem.Allocator,
    nodes: std.ArrayList(rtb.RtbBvhNode),
    hittables: std.ArrayList(rtb.RtbHittable),
    materials: std.ArrayList(rtb.RtbMaterial),
    textures: std.ArrayList(rtb.RtbTexture),
    perlins: std.ArrayList(rtb.RtbPerlin),
    images: std.ArrayList(rtb.RtbImage),
    root: i32 = -1,

    pub fn init(allocator: std.mem.Allocator) Lowered {
        return .{
            .allocator = allocator,
            .nodes = std.ArrayList(rtb.RtbBvhNode).init(allocator),
            .hittables = std.ArrayList(rtb.RtbHittable).init(allocator),
            .materials = std.ArrayList(rtb.RtbMaterial).init(allocator),
            .textures = std.ArrayList(rtb.RtbTexture).init(allocator),
            .perlins = std.ArrayList(rtb.RtbPerlin).init(allocator),
            .images = std.ArrayList(rtb.RtbImage).init(allocator),
        };
    }

    pub fn deinit(self: *Lowered) void {
        self.nodes.deinit();
        self.hittables.deinit();
        self.materials.deinit();
        self.textures.deinit();
        self.perlins.deinit();
        self.images.deinit();
    }

    pub fn desc(self: *const Lowered) rtb.RtbSceneDesc {
        return .{
            .n_nodes = @intCast(self.nodes.items.len),
            .n_hittables = @intCast(self.hittables.items.len),
            .n_materials = @intCast(self.materials.items.len),
            .n_textures = @intCast(self.textures.items.len),
            .n_perlins = @intCast(self.perlins.items.len),
            .n_images = @intCast(self.images.items.len),
            .root = self.root,
            .nodes = self.nodes.items.ptr,
            .hittables = self.hittables.items.ptr,
            .materials = self.materials.items.ptr,
            .textures = self.textures.items.ptr,
            .perlins = self.perlins.items.ptr,
            .images = self.images.items.ptr,
        };
    }
};

/// Texture union (src/textures.zig:10-14) -> RtbTexture; returns its index.
pub fn lowerTexture(l: *Lowered, tex: Texture) LowerError!u32 {
    const index: u32 = @intCast(l.textures.items.len);
    switch (tex) {
        .solid_color => |s| try l.textures.append(.{ .type = rtb.TEX_SOLID, .color = v3(s.color_value) }),
        // CheckerTexture keeps inv_scale = 1 / scale (src/textures.zig:53-56); even / odd are SolidColor only
        .checker_texture => |c| try l.textures.append(.{ .type = rtb.TEX_CHECKER, .scale = c.inv_scale, .color = v3(c.even.color_value), .color2 = v3(c.odd.color_value) }),
        // ImageTexture -> RtwImage{ images, image_index } (src/rtw_image.zig:5-7): the texel data is lowered by lowerImages
        .image_texture => |i| try l.textures.append(.{ .type = rtb.TEX_IMAGE, .index = i.rtw_image.image_index }),
        // NoiseTexture owns its Perlin tables BY VALUE (src/textures.zig:106-109, src/perlin.zig:76-81): copy them
        .noise_texture => |n| {
            var p: rtb.RtbPerlin = undefined;
            for (0..256) |k| {
                p.ranvec[k] = v3(n.noise.ranvec[k]);
                p.perm_x[k] = n.noise.perm_x[k];
                p.perm_y[k] = n.noise.perm_y[k];
                p.perm_z[k] = n.noise.perm_z[k];
            }
            const slot: u32 = @intCast(l.perlins.items.len);
            try l.perlins.append(p);
            try l.textures.append(.{ .type = rtb.TEX_NOISE, .index = slot, .scale = n.scale });
        },
    }
    return index;
}

/// Material union (src/material.zig:11-16) -> RtbMaterial; returns its index.  The reference stores Material by value
/// in every Hittable, so one record per object is emitted (no attempt to share).
pub fn lowerMaterial(l: *Lowered, mat: Material) LowerError!u32 {
    const index: u32 = @intCast(l.materials.items.len);
    switch (mat) {
        .lambertian => |m| {
            const t = try lowerTexture(l, m.albedo);
            try l.materials.append(.{ .type = rtb.MAT_LAMBERTIAN, .texture = t });
        },
        .metal => |m| try l.materials.append(.{ .type = rtb.MAT_METAL, .albedo = v3(m.albedo), .fuzz = m.fuzz }), // fuzz already clamped (:61-63)
        .dielectric => |m| try l.materials.append(.{ .type = rtb.MAT_DIELECTRIC, .ir = m.ir }),
        .diffuse_light => |m| {
            const t = try lowerTexture(l, m.emit);
            try l.materials.append(.{ .type = rtb.MAT_DIFFUSE_LIGHT, .texture = t });
        },
        .isotropic => |m| {
            const t = try lowerTexture(l, m.albedo);
            try l.materials.append(.{ .type = rtb.MAT_ISOTROPIC, .texture = t });
        },
    }
    // lowerTexture may have appended after `index` was taken only for textures, never for materials: index is right
    return index;
}

/// A `createBox(a, b, mat)` list (src/objects.zig:510-532): six quads that share one material.  The two corners handed
/// to createBox are recovered as the extrema of the quads' corners (q, q + u, q + v, q + u + v); the library rebuilds
/// the same six quads from them, in createBox's order — including HEAD's quirk of emitting the z = min face twice and
/// no z = max face (:520-529), which tests/test_cornell.py pins.  Returns false if the list is not such a box.
fn boxCorners(list: objects.HittableList, a: *[3]f32, b: *[3]f32, mat: *Material) bool {
    if (list.objects.items.len != 6) return false;
    var mn = [3]f32{ std.math.inf(f32), std.math.inf(f32), std.math.inf(f32) };
    var mx = [3]f32{ -std.math.inf(f32), -std.math.inf(f32), -std.math.inf(f32) };
    for (list.objects.items) |side| {
        switch (side) {
            .quad => |q| {
                // createBox's quads are axis-aligned: u and v each have exactly one non-zero component
                var nz: u32 = 0;
                for (0..3) |k| {
                    if (q.u[k] != 0) nz += 1;
                    if (q.v[k] != 0) nz += 1;
                }
                if (nz != 2) return false;
                const corners = [4]Vec3{ q.q, q.q + q.u, q.q + q.v, q.q + q.u + q.v };
                for (corners) |c| {
                    for (0..3) |k| {
                        mn[k] = @min(mn[k], c[k]);
                        mx[k] = @max(mx[k], c[k]);
                    }
                }
                mat.* = q.mat;
            },
            else => return false,
        }
    }
    a.* = mn;
    b.* = mx;
    return true;
}

/// The fast path for the instancing shape HEAD's scenes use (src/main.zig:182-190, :223-236): a createBox list,
/// optionally wrapped in RotateY and / or Translate (in that nesting), becomes ONE RtbHittable of type BOX.
/// Returns null when `h` is anything else — the caller then lowers it with the general wrappers.
fn lowerBoxInstance(h: Hittable, out: *rtb.RtbHittable) ?Material {
    var cur = h;
    out.c = .{ 0, 0, 0 };
    out.sin_theta = 0;
    out.cos_theta = 1;
    if (cur == .translate) { // Translate.offset (src/objects.zig:309)
        out.c = v3(cur.translate.offset);
        cur = cur.translate.object.*;
    }
    if (cur == .rotate_y) { // RotateY.sin_theta / cos_theta (src/objects.zig:350-358)
        out.sin_theta = cur.rotate_y.sin_theta;
        out.cos_theta = cur.rotate_y.cos_theta;
        cur = cur.rotate_y.object.*;
    }
    if (cur != .list) return null;
    var mat: Material = undefined;
    if (!boxCorners(cur.list, &out.a, &out.b, &mat)) return null;
    return mat;
}

/// Any Hittable (src/objects.zig:39-47) -> l.hittables[at].  Wrappers (Translate / RotateY / HittableList /
/// ConstantMedium over anything) refer to their wrapped objects through `child`; those are appended AFTER everything
/// already in the array (a child always follows its wrapper, list members are contiguous), recursively.
pub fn lowerInto(l: *Lowered, at: usize, h: Hittable) LowerError!void {
    var out = rtb.RtbHittable{ .type = rtb.HITTABLE_SPHERE, .material = 0 };
    switch (h) {
        .sphere => |s| {
            out.type = rtb.HITTABLE_SPHERE;
            out.a = v3(s.center1);
            out.b = v3(s.center_vec); // zero unless is_moving (src/objects.zig:87-92)
            out.radius = s.radius;
            out.is_moving = @intFromBool(s.is_moving);
            out.material = try lowerMaterial(l, s.mat);
        },
        .quad => |q| {
            out.type = rtb.HITTABLE_QUAD;
            out.a = v3(q.q);
            out.b = v3(q.u);
            out.c = v3(q.v);
            out.material = try lowerMaterial(l, q.mat);
        },
        .list, .translate, .rotate_y => {
            if (lowerBoxInstance(h, &out)) |mat| { // the one-record fast path
                out.type = rtb.HITTABLE_BOX;
                out.material = try lowerMaterial(l, mat);
            } else switch (h) {
                .translate => |t| { // Translate{offset, object} (src/objects.zig:308-311)
                    out.type = rtb.HITTABLE_TRANSLATE;
                    out.a = v3(t.offset);
                    out.child = @intCast(l.hittables.items.len);
                    try l.hittables.append(undefined);
                    try lowerInto(l, out.child, t.object.*);
                },
                .rotate_y => |r| { // RotateY{object, sin_theta, cos_theta} (:348-352)
                    out.type = rtb.HITTABLE_ROTATE_Y;
                    out.sin_theta = r.sin_theta;
                    out.cos_theta = r.cos_theta;
                    out.child = @intCast(l.hittables.items.len);
                    try l.hittables.append(undefined);
                    try lowerInto(l, out.child, r.object.*);
                },
                .list => |ls| { // HittableList{objects} (:264-267): members contiguous, in order
                    if (ls.objects.items.len == 0) return LowerError.Unsupported;
                    out.type = rtb.HITTABLE_LIST;
                    out.material = @intCast(ls.objects.items.len); // a list has no material: the field carries the count
                    out.child = @intCast(l.hittables.items.len);
                    for (ls.objects.items) |_| try l.hittables.append(undefined);
                    for (ls.objects.items, 0..) |member, k| try lowerInto(l, out.child + k, member);
                },
                else => unreachable,
            }
        },
        .constant_medium => |m| { // src/objects.zig:445-452
            out.radius = m.neg_inv_density; // -1 / density (:451)
            out.material = try lowerMaterial(l, m.phase_function);
            if (lowerBoxInstance(m.boundary.*, &out)) |_| {
                out.type = rtb.HITTABLE_CONSTANT_MEDIUM; // fog in a box instance: one record (cornellBoxSmoke)
            } else {
                out.type = rtb.HITTABLE_MEDIUM_OF; // fog inside any other boundary
                out.child = @intCast(l.hittables.items.len);
                try l.hittables.append(undefined);
                try lowerInto(l, out.child, m.boundary.*);
            }
        },
        .round_box => return LowerError.Unsupported, // unfinished in the reference (src/objects.zig:171-192)
        .tree => return LowerError.Unsupported, // a tree inside a tree does not occur in HEAD's scenes
    }
    l.hittables.items[at] = out;
}

/// One `world_objects.items[i]` -> RtbHittable at the SAME index i: the position in the object list is the "object
/// index" rtb_trace_rays reports, and BVH leaves refer to it.  (Children of wrappers land behind the list.)
pub fn lowerHittable(l: *Lowered, i: usize, h: Hittable) LowerError!void {
    try lowerInto(l, i, h);
}

/// BVHNode pointer graph (src/bvh.zig:106-110) -> index-linked RtbBvhNode array, any order (the library re-lays the
/// tree out for the device).  `base` = world_objects.items.ptr: leaves point INTO that slice (src/bvh.zig:51-57), so the
/// object index is the pointer difference.
pub fn lowerNode(l: *Lowered, n: *const BVHNode, base: [*]const Hittable) LowerError!i32 {
    const me: usize = l.nodes.items.len;
    const bb = n.bounding_box;
    try l.nodes.append(.{ .bmin = .{ bb.x.min, bb.y.min, bb.z.min }, .bmax = .{ bb.x.max, bb.y.max, bb.z.max } });
    if (n.leaf) |h| {
        l.nodes.items[me].leaf = @intCast((@intFromPtr(h) - @intFromPtr(base)) / @sizeOf(Hittable));
    } else {
        const left = try lowerNode(l, n.left.?, base);
        const right = try lowerNode(l, n.right.?, base);
        l.nodes.items[me].left = left;
        l.nodes.items[me].right = right;
    }
    return @intCast(me);
}

/// zstbi images as loaded by `zstbi.Image.loadFromFile(path, 4)` (src/main.zig:1124): RGBA8, row stride bytes_per_row.
pub fn lowerImages(l: *Lowered, images: std.ArrayList(zstbi.Image)) LowerError!void {
    for (images.items) |im| {
        try l.images.append(.{ .width = im.width, .height = im.height, .bytes_per_row = im.bytes_per_row, .data = im.data.ptr });
    }
}

/// world = Hittable{ .tree = BVHTree.init(allocator, world_objects.items, 0, len) } (e.g. src/main.zig:309-311).
pub fn lowerWorld(allocator: std.mem.Allocator, world: Hittable, world_objects: []const Hittable, images: std.ArrayList(zstbi.Image)) LowerError!Lowered {
    var l = Lowered.init(allocator);
    errdefer l.deinit();
    try l.hittables.resize(world_objects.len); // index i of world_objects.items -> hittables[i]; children follow
    for (world_objects, 0..) |h, i| try lowerHittable(&l, i, h);
    try lowerImages(&l, images);
    switch (world) {
        .tree => |t| l.root = try lowerNode(&l, t.root, world_objects.ptr),
        else => return LowerError.Unsupported,
    }
    return l;
}

/// Camera after `camera.init()` (src/camera.zig:118-154) -> RtbCamera: only the derived fields the hot loop reads.
pub fn lowerCamera(cam: Camera) rtb.RtbCamera {
    return .{
        .image_width = cam.image_width,
        .image_height = cam.image_height,
        .samples_per_pixel = cam.samples_per_pixel,
        .max_depth = cam.max_depth,
        .center = v3(cam.center),
        .pixel00_loc = v3(cam.pixel00_loc),
        .pixel_delta_u = v3(cam.pixel_delta_u),
        .pixel_delta_v = v3(cam.pixel_delta_v),
        .defocus_disk_u = v3(cam.defocus_disk_u),
        .defocus_disk_v = v3(cam.defocus_disk_v),
        .defocus_angle = cam.defocus_angle,
        .background = v3(cam.background),
        .background_mode = rtb.BACKGROUND_SOLID,
    };
}

// ------------------------------------------------------------------------------------------------------------------
// Replacing the render threads.  RayTraceState (src/main.zig:71-86) gains three fields:
//     scene: ?*rtb.RtbScene = null,  job: ?*rtb.RtbJob = null,  group: ?*rtb.RtbSceneGroup = null,
// and `threads` / RenderThread (src/main.zig:49-69) go away.
// ------------------------------------------------------------------------------------------------------------------

/// Call once after the scene builder returned `world` (src/main.zig:422-430): uploads the scene to GPU `device`.
pub fn uploadScene(allocator: std.mem.Allocator, world: Hittable, world_objects: []const Hittable, images: std.ArrayList(zstbi.Image), device: c_int) !*rtb.RtbScene {
    var l = try lowerWorld(allocator, world, world_objects, images);
    defer l.deinit(); // the library has copied everything when rtb_scene_create returns
    const d = l.desc();
    var scene: ?*rtb.RtbScene = null;
    if (rtb.rtb_scene_create(&d, device, &scene) != rtb.OK) {
        std.log.err("rtb_scene_create: {s}", .{rtb.rtb_last_error()});
        return error.RtbSceneCreateFailed;
    }
    return scene.?;
}

/// startRender (src/main.zig:314-326): one asynchronous job instead of 8 threads over pixel strips.  The job's worker
/// refreshes writer.buffer / writer.texture_buffer after every sample, exactly what writeColor does per pixel-sample
/// (src/camera.zig:54-66), so countSamples (:470-477) and updateTexture (:568-612) keep working unchanged.
pub fn startRender(raytrace: anytype) !void {
    raytrace.writer.scrub();
    raytrace.render_running.* = true;
    try raytrace.camera.init(); // reinitialise with the GUI's parameters, as before
    const cam = lowerCamera(raytrace.camera.*);
    const opt = rtb.RtbRenderOptions{ .samples_per_launch = 1, .traversal = rtb.TRAVERSAL_SAH16 };
    var job: ?*rtb.RtbJob = null;
    if (rtb.rtb_render_async(raytrace.scene.?, &cam, &opt, @ptrCast(raytrace.writer.buffer.ptr), raytrace.writer.texture_buffer.ptr, &job) != rtb.OK) {
        std.log.err("rtb_render_async: {s}", .{rtb.rtb_last_error()});
        return error.RtbRenderFailed;
    }
    raytrace.job = job;
    raytrace.render_start.* = std.time.milliTimestamp();
}

/// stopRender (src/main.zig:328-336): the STOP button.  Cancel takes effect at the next sample boundary.
pub fn stopRender(raytrace: anytype) !void {
    raytrace.render_running.* = false;
    raytrace.render_end.* = std.time.milliTimestamp();
    if (raytrace.job) |job| {
        _ = rtb.rtb_job_cancel(job);
        _ = rtb.rtb_job_wait(job, null); // RTB_ERR_CANCELLED is the expected status after a STOP
        _ = rtb.rtb_job_destroy(job);
        raytrace.job = null;
    }
}

/// shouldStopRender (src/main.zig:338-348): polled once per GUI frame.
pub fn shouldStopRender(raytrace: anytype) !void {
    if (raytrace.job) |job| {
        var running: c_int = 0;
        _ = rtb.rtb_job_progress(job, null, null, &running);
        if (running == 0 and raytrace.render_running.*) try stopRender(raytrace);
    }
}

/// All GPUs of the box behind one blocking call (no progressive display): the frame ends up in writer.buffer /
/// writer.texture_buffer like after the 8 threads have joined.
pub fn renderOnAllGpus(allocator: std.mem.Allocator, raytrace: anytype, world_objects: []const Hittable) !void {
    var n: c_int = 0;
    if (rtb.rtb_device_count(&n) != rtb.OK) return error.NoCudaDevice;
    if (raytrace.group == null) {
        var l = try lowerWorld(allocator, raytrace.world, world_objects, raytrace.images);
        defer l.deinit();
        const d = l.desc();
        const devices = try allocator.alloc(c_int, @intCast(n));
        defer allocator.free(devices);
        for (devices, 0..) |*dev, k| dev.* = @intCast(k);
        var group: ?*rtb.RtbSceneGroup = null;
        if (rtb.rtb_group_create(&d, devices.ptr, @intCast(n), &group) != rtb.OK) return error.RtbGroupCreateFailed;
        raytrace.group = group;
    }
    raytrace.writer.scrub();
    try raytrace.camera.init();
    const cam = lowerCamera(raytrace.camera.*);
    const opt = rtb.RtbRenderOptions{ .traversal = rtb.TRAVERSAL_SAH16 };
    if (rtb.rtb_group_render(raytrace.group.?, &cam, &opt, rtb.PARTITION_SAMPLES, @ptrCast(raytrace.writer.buffer.ptr), raytrace.writer.texture_buffer.ptr, null) != rtb.OK) {
        std.log.err("rtb_group_render: {s}", .{rtb.rtb_last_error()});
        return error.RtbRenderFailed;
    }
}
