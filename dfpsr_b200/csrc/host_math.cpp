// host_math.cpp — host arithmetic of the C ABI that is hot enough to vectorise. Compiled by the host compiler only (no CUDA front end), with
// -ffp-contract=off: every lane performs the reference's IEEE operations in the reference's order, so results are bit-identical to scalar code.
#include "../../include/dfpsr_b200.h"

typedef float v4f __attribute__((vector_size(16)));
typedef int v4i __attribute__((vector_size(16)));

static inline v4f splat(float v) { return v4f{v, v, v, v}; }

// ref: implementation/render/Camera.h:202-217 Camera::isBoxSeen, ViewFrustum :41-109 — 0 hidden, 1 partial, 2 fully inside.
// The eight corners (corner i takes max on axis k when bit k of i is set) go through modelToWorld.transformPoint (math/Transform3D.h:41-43)
// and worldToCamera (:50-52) four at a time. A Sandbox frame asks this for about a thousand shadow submissions, which took longer than the
// kernels they feed.
extern "C" int dfpsr_camera_is_box_seen(const dfpsr_camera *c, const float mn[3], const float mx[3], const dfpsr_transform3d *m2w) {
	const v4f px = {mn[0], mx[0], mn[0], mx[0]}, py = {mn[1], mn[1], mx[1], mx[1]};
	const v4f pz[2] = {splat(mn[2]), splat(mx[2])};
	const dfpsr_transform3d &l = c->location;
	v4f cx[2], cy[2], cz[2];
	for (int h = 0; h < 2; h++) {
		const v4f wx = (px * splat(m2w->xAxis[0]) + py * splat(m2w->yAxis[0]) + pz[h] * splat(m2w->zAxis[0])) + splat(m2w->position[0]);
		const v4f wy = (px * splat(m2w->xAxis[1]) + py * splat(m2w->yAxis[1]) + pz[h] * splat(m2w->zAxis[1])) + splat(m2w->position[1]);
		const v4f wz = (px * splat(m2w->xAxis[2]) + py * splat(m2w->yAxis[2]) + pz[h] * splat(m2w->zAxis[2])) + splat(m2w->position[2]);
		const v4f dx = wx - splat(l.position[0]), dy = wy - splat(l.position[1]), dz = wz - splat(l.position[2]);
		cx[h] = dx * splat(l.xAxis[0]) + dy * splat(l.xAxis[1]) + dz * splat(l.xAxis[2]);
		cy[h] = dx * splat(l.yAxis[0]) + dy * splat(l.yAxis[1]) + dz * splat(l.yAxis[2]);
		cz[h] = dx * splat(l.zAxis[0]) + dy * splat(l.zAxis[1]) + dz * splat(l.zAxis[2]);
	}
	bool anyOutside = false;
	for (int s = 0; s < c->cullPlaneCount; s++) {
		const float *pl = c->cullPlanes[s];
		const v4f nx = splat(pl[0]), ny = splat(pl[1]), nz = splat(pl[2]), off = splat(pl[3]);
		const v4f d0 = ((nx * cx[0]) + (ny * cy[0]) + (nz * cz[0])) - off, d1 = ((nx * cx[1]) + (ny * cy[1]) + (nz * cz[1])) - off;
		const v4i in0 = d0 <= splat(0.0f), in1 = d1 <= splat(0.0f); // -1 where the corner is inside (a NaN distance counts as outside)
		const v4i any = in0 | in1, all = in0 & in1;
		if (!(any[0] | any[1] | any[2] | any[3])) { return 0; }
		if (!(all[0] & all[1] & all[2] & all[3])) { anyOutside = true; }
	}
	return anyOutside ? 1 : 2;
}
