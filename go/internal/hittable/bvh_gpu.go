// STATUS: source only (no Go toolchain in this repository's build environment; see INTEGRATION.md).
//
// BuildBVH with the sorts on the GPU.  bvhHelper (bvh.go:35-61) always splits at the median, so the shape of the
// tree follows from len(list.objects) alone; what depends on the data is the order the objects end up in after the
// recursive sort.Slice calls.  libgrt_cuda computes that order (grt_bvh_order, include/grt.h); the nodes are then
// built over the permuted slice without sorting.  Package cuda installs the hook (it imports this package, so the
// dependency cannot point the other way).
package hittable

import "github.com/nsp5488/go_raytracer/internal/aabb"

// BVHOrderHook, when set, returns the final object order for the n boxes {lo.x,lo.y,lo.z,hi.x,hi.y,hi.z}.
var BVHOrderHook func(boxes []float64) ([]uint32, error)

// BVHOrderMin is the list length from which the hook is used (below it the host sort is faster than the round trip).
var BVHOrderMin = 32768

// BuildBVHFast is BuildBVH (bvh.go:21-23) with the object order taken from the GPU when a hook is installed.
// Ties (equal box minimum and maximum on the split axis) keep list order; sort.Slice leaves them unspecified.
func BuildBVHFast(list *HittableList) *BVHNode {
	n := len(list.objects)
	if BVHOrderHook == nil || n < BVHOrderMin {
		return BuildBVH(list)
	}
	boxes := make([]float64, 6*n)
	for i, o := range list.objects {
		b := o.BBox()
		for a := 0; a < 3; a++ {
			iv := b.AxisInterval(a)
			boxes[6*i+a], boxes[6*i+3+a] = iv.Min, iv.Max
		}
	}
	order, err := BVHOrderHook(boxes)
	if err != nil {
		return BuildBVH(list)
	}
	sorted := make([]Hittable, n)
	for p, idx := range order {
		sorted[p] = list.objects[idx]
	}
	copy(list.objects, sorted)
	return presortedHelper(list, 0, n)
}

// presortedHelper is bvhHelper without the sort: the slice already stands in its final order, and a span's box is the
// union of its halves' boxes.
func presortedHelper(list *HittableList, start, end int) *BVHNode {
	span := end - start
	switch {
	case span == 1:
		o := list.objects[start]
		return &BVHNode{left: o, right: o, bbox: o.BBox()}
	case span == 2:
		l, r := list.objects[start], list.objects[start+1]
		return &BVHNode{left: l, right: r, bbox: aabb.FromBBoxes(l.BBox(), r.BBox())}
	}
	mid := start + span/2
	l, r := presortedHelper(list, start, mid), presortedHelper(list, mid, end)
	return &BVHNode{left: l, right: r, bbox: aabb.FromBBoxes(l.bbox, r.bbox)}
}
