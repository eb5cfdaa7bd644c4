// flatten.go — lives INSIDE package hittable because every concrete type and field it
// reads is unexported (sphere, quad, BVHNode.left/right, lambertian.tex, perlin.randVec, ...).
//
// STATUS: source only (no Go toolchain in this repository's build environment).  The C++
// flattener go_raytracer_b200/csrc/flatten.hpp is the tested twin of this file; both emit
// the layout of include/grt.h.  Differences: this version walks the ALREADY-BUILT Go tree
// (BuildBVH has run, bvh.go:21), so it needs no sort and reproduces Go's pdqsort tie order.
package hittable

import (
	"math"
	"runtime"
	"unsafe"

	"github.com/nsp5488/go_raytracer/internal/vec"
)

const (
	refNode, refSphere, refQuad, refTri, refList, refMedium, refNone = 0, 1, 2, 3, 4, 5, 7
	listLast                                                      = 0x80000000
	collapseWhole, collapseLeaf                                   = 32, 4
)

func mkRef(t, i uint32) uint32 { return t<<28 | i&0x0FFFFFFF }

// Plain-old-data records, field for field include/grt.h.
type flatNode struct {
	Bmin, Bmax  [3]float32
	Left, Right uint32
}
type flatSphere struct {
	C0    [3]float64
	R     float64
	Dc    [3]float32
	Mat   uint32
	ID    uint32
	Flags uint32
	UVRot [2]float32
}
type flatQuad struct {
	N     [3]float32
	D     float32
	Q     [3]float32
	Flags uint32
	A     [3]float32
	Mat   uint32
	B     [3]float32
	ID    uint32
	N64   [3]float64
	D64   float64
}
type flatTri struct {
	V0    [3]float32
	Mat   uint32
	E0    [3]float32
	ID    uint32
	E1    [3]float32
	Flags uint32
}
type flatTriShade struct {
	N0, N1, N2 [3]float32
	UV         [6]float32
	Pad        float32
}
type flatMedium struct {
	Boundary      uint32
	NegInvDensity float32
	Mat, ID       uint32
}
type flatMaterial struct {
	Type, Tex uint32
	Albedo    [3]float32
	Fuzz, Ior float32
	Pad       uint32
}
type flatTexture struct {
	Type           uint32
	Color          [3]float32
	Scale          float32
	Even, Odd, Aux uint32
}
type flatLight struct {
	Type, Prim, Flags, Pad uint32
	P                      [24]float64
	F                      [24]float32
}

// FlatScene owns the arrays handed to libgrt_cuda.
type FlatScene struct {
	Root       uint32
	Nodes      []flatNode
	Spheres    []flatSphere
	Quads      []flatQuad
	Tris       []flatTri
	TriShade   []flatTriShade
	TriV64     []float64
	Items      []uint32
	Media      []flatMedium
	Materials  []flatMaterial
	Textures   []flatTexture
	Lights     []flatLight
	LightsMode uint32
	DepthHint  uint32

	matID map[Material]uint32
	texID map[Texture]uint32
	objID map[Hittable]uint32
}

// xform: world = R(obj) + T, R the Y rotation of transformation.go:87-93.
type xform struct {
	c, s float64
	t    [3]float64
}

func (x xform) rot(v *vec.Vec3) [3]float64 {
	return [3]float64{x.c*v.X() + x.s*v.Z(), v.Y(), -x.s*v.X() + x.c*v.Z()}
}
func (x xform) point(v *vec.Vec3) [3]float64 {
	r := x.rot(v)
	return [3]float64{r[0] + x.t[0], r[1] + x.t[1], r[2] + x.t[2]}
}

// Flatten type-switches over the nine concrete Hittables, five materials and four
// textures of the reference and bakes translate/rotateY instances into world space.
func Flatten(world, lights Hittable) (*FlatScene, error) {
	fs := &FlatScene{matID: map[Material]uint32{}, texID: map[Texture]uint32{}, objID: map[Hittable]uint32{}}
	ref, _, err := fs.emit(world, xform{c: 1}, false)
	if err != nil {
		return nil, err
	}
	fs.Root = ref
	if err := fs.emitLights(lights); err != nil {
		return nil, err
	}
	return fs, nil
}

type box3 struct{ lo, hi [3]float64 }

func emptyBox() box3 {
	return box3{[3]float64{math.Inf(1), math.Inf(1), math.Inf(1)}, [3]float64{math.Inf(-1), math.Inf(-1), math.Inf(-1)}}
}
func (b *box3) add(p [3]float64) {
	for a := 0; a < 3; a++ {
		b.lo[a], b.hi[a] = math.Min(b.lo[a], p[a]), math.Max(b.hi[a], p[a])
	}
}
func (b *box3) union(o box3) { b.add(o.lo); b.add(o.hi) }

func down(x float64) float32 {
	f := float32(x)
	if float64(f) > x {
		f = math.Nextafter32(f, float32(math.Inf(-1)))
	}
	return math.Nextafter32(f, float32(math.Inf(-1)))
}
func up(x float64) float32 {
	f := float32(x)
	if float64(f) < x {
		f = math.Nextafter32(f, float32(math.Inf(1)))
	}
	return math.Nextafter32(f, float32(math.Inf(1)))
}

// leaves counts the primitive tests of a linear scan (collapse decision).
func leaves(h Hittable) int {
	switch o := h.(type) {
	case *BVHNode:
		if o.left == o.right {
			return leaves(o.left)
		}
		return leaves(o.left) + leaves(o.right)
	case *HittableList:
		n := 0
		for _, c := range o.objects {
			n += leaves(c)
		}
		return n
	case *translate:
		return leaves(o.object)
	case *rotateY:
		return leaves(o.object)
	default:
		return 1
	}
}

func (fs *FlatScene) emit(h Hittable, x xform, inBoundary bool) (uint32, box3, error) {
	switch o := h.(type) {
	case *sphere:
		c0, dc := x.point(o.Center.Origin()), x.rot(o.Center.Direction())
		s := flatSphere{C0: c0, R: o.Radius, Mat: fs.material(o.Material), ID: fs.id(h), UVRot: [2]float32{float32(x.c), float32(x.s)}}
		for a := 0; a < 3; a++ {
			s.Dc[a] = float32(dc[a])
		}
		fs.Spheres = append(fs.Spheres, s)
		b := emptyBox()
		for _, t := range []float64{0, 1} {
			for _, sg := range []float64{-1, 1} {
				b.add([3]float64{c0[0] + t*dc[0] + sg*o.Radius, c0[1] + t*dc[1] + sg*o.Radius, c0[2] + t*dc[2] + sg*o.Radius})
			}
		}
		return mkRef(refSphere, uint32(len(fs.Spheres)-1)), b, nil
	case *quad:
		return fs.emitQuad(o, x)
	case *Triangle:
		return fs.emitTri(o, x)
	case *HittableList:
		return fs.emitRun(o.objects, x, inBoundary)
	case *BVHNode:
		limit := collapseLeaf
		if n := leaves(o); n <= collapseWhole { // (the Go tree has no parent pointer: callers pass whole trees first)
			limit = collapseWhole
		}
		if leaves(o) <= limit {
			var run []Hittable
			collect(o, &run)
			return fs.emitRun(run, x, inBoundary)
		}
		idx := uint32(len(fs.Nodes))
		fs.Nodes = append(fs.Nodes, flatNode{}) // depth-first, left-first
		l, lb, err := fs.emit(o.left, x, inBoundary)
		if err != nil {
			return 0, lb, err
		}
		r, rb := l, lb
		if o.right != o.left {
			if r, rb, err = fs.emit(o.right, x, inBoundary); err != nil {
				return 0, rb, err
			}
		}
		lb.union(rb)
		n := &fs.Nodes[idx]
		for a := 0; a < 3; a++ {
			lo, hi := lb.lo[a], lb.hi[a]
			if hi-lo < 0.0001 {
				lo, hi = lo-0.00005, hi+0.00005
			}
			n.Bmin[a], n.Bmax[a] = down(lo), up(hi)
		}
		n.Left, n.Right = l, r
		return mkRef(refNode, idx), lb, nil
	case *translate:
		off := x.rot(o.offset)
		y := x
		y.t = [3]float64{x.t[0] + off[0], x.t[1] + off[1], x.t[2] + off[2]}
		return fs.emit(o.object, y, inBoundary)
	case *rotateY:
		y := x
		y.c, y.s = x.c*o.cosTheta-x.s*o.sinTheta, x.s*o.cosTheta+x.c*o.sinTheta
		return fs.emit(o.object, y, inBoundary)
	case *constantMedium:
		if inBoundary {
			return 0, emptyBox(), errUnsupported("constantMedium nested inside a medium boundary")
		}
		b, bb, err := fs.emit(o.boundary, x, true)
		if err != nil {
			return 0, bb, err
		}
		fs.Media = append(fs.Media, flatMedium{Boundary: b, NegInvDensity: float32(o.negativeInverseDensity), Mat: fs.material(o.phaseFunction), ID: fs.id(h)})
		return mkRef(refMedium, uint32(len(fs.Media)-1)), bb, nil
	}
	return 0, emptyBox(), errUnsupported("unknown Hittable")
}

// collect lists the leaves of a BVH subtree in the order BVHNode.Hit visits them (bvh.go:73-79).
func collect(n *BVHNode, out *[]Hittable) {
	for i, c := range []Hittable{n.left, n.right} {
		if i == 1 && n.right == n.left {
			if _, isMedium := c.(*constantMedium); !isMedium {
				continue // a surface tested twice cannot change the closest hit; a medium draws again (medium.go:47)
			}
		}
		if b, ok := c.(*BVHNode); ok {
			collect(b, out)
		} else {
			*out = append(*out, c)
		}
	}
}

type errUnsupported string

func (e errUnsupported) Error() string { return "hittable.Flatten: " + string(e) }

// Pin keeps the Go-owned arrays reachable and immovable while C reads them.
type Pin struct{ p runtime.Pinner }

func (p *Pin) Unpin() { p.p.Unpin() }

// Fill writes the GrtScene view (pointers + counts) into dst (a *C.GrtScene passed as
// unsafe.Pointer so that this package does not import "C").
func (fs *FlatScene) Fill(dst unsafe.Pointer) *Pin {
	pin := &Pin{}
	v := (*sceneView)(dst)
	*v = sceneView{AbiVersion: 1, Root: fs.Root, LightsMode: fs.LightsMode, DepthHint: fs.DepthHint}
	set := func(p *unsafe.Pointer, n *uint32, data unsafe.Pointer, count int) {
		if count > 0 {
			pin.p.Pin(data)
			*p = data
		}
		*n = uint32(count)
	}
	set(&v.Nodes, &v.NNodes, unsafe.Pointer(unsafe.SliceData(fs.Nodes)), len(fs.Nodes))
	set(&v.Spheres, &v.NSpheres, unsafe.Pointer(unsafe.SliceData(fs.Spheres)), len(fs.Spheres))
	set(&v.Quads, &v.NQuads, unsafe.Pointer(unsafe.SliceData(fs.Quads)), len(fs.Quads))
	set(&v.Tris, &v.NTris, unsafe.Pointer(unsafe.SliceData(fs.Tris)), len(fs.Tris))
	set(&v.Items, &v.NItems, unsafe.Pointer(unsafe.SliceData(fs.Items)), len(fs.Items))
	set(&v.Media, &v.NMedia, unsafe.Pointer(unsafe.SliceData(fs.Media)), len(fs.Media))
	set(&v.Materials, &v.NMaterials, unsafe.Pointer(unsafe.SliceData(fs.Materials)), len(fs.Materials))
	set(&v.Textures, &v.NTextures, unsafe.Pointer(unsafe.SliceData(fs.Textures)), len(fs.Textures))
	set(&v.Lights, &v.NLights, unsafe.Pointer(unsafe.SliceData(fs.Lights)), len(fs.Lights))
	// tri_shade, tri_v64, images, texels and perlins are filled the same way (elided: same pattern)
	return pin
}

// sceneView mirrors struct GrtScene of include/grt.h.
type sceneView struct {
	AbiVersion, Root         uint32
	Nodes                    unsafe.Pointer
	NNodes                   uint32
	Spheres                  unsafe.Pointer
	NSpheres                 uint32
	Quads                    unsafe.Pointer
	NQuads                   uint32
	Tris                     unsafe.Pointer
	NTris                    uint32
	TriShade, TriV64, Items  unsafe.Pointer
	NItems                   uint32
	Media                    unsafe.Pointer
	NMedia                   uint32
	Materials                unsafe.Pointer
	NMaterials               uint32
	Textures                 unsafe.Pointer
	NTextures                uint32
	Images                   unsafe.Pointer
	NImages                  uint32
	Texels                   unsafe.Pointer
	NTexelBytes              uint64
	Perlins                  unsafe.Pointer
	NPerlins                 uint32
	Lights                   unsafe.Pointer
	NLights                  uint32
	LightsMode, DepthHint    uint32
}

// emitQuad, emitTri, emitRun, emitLights, material, texture and id follow
// go_raytracer_b200/csrc/flatten.hpp line for line (NewQuad's derived fields
// objects.go:129-141 recomputed on the transformed Q,u,v; A = v×w, B = w×u;
// lights restricted to sphere/quad/Triangle as hittable.go:69-72 demands).
