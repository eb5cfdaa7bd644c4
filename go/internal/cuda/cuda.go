// Package cuda binds libgrt_cuda (include/grt.h) for go_raytracer.
//
// STATUS: source only.  The build environment of this repository has no Go
// toolchain, so this file has never been compiled; it documents the exact
// binding a maintainer adds to the reference (see INTEGRATION.md).  The same
// C entry points are exercised from C++ (csrc/main.cpp) and Python (ctypes).
package cuda

/*
#cgo CFLAGS: -I${SRCDIR}/../../../include
#cgo LDFLAGS: -L${SRCDIR}/../../../go_raytracer_b200/csrc -lgrt_cuda -Wl,-rpath,${SRCDIR}/../../../go_raytracer_b200/csrc
#include <stdlib.h>
#include "grt.h"
*/
import "C"

import (
	"errors"
	"fmt"
	"io"
	"runtime"
	"unsafe"

	"github.com/nsp5488/go_raytracer/internal/hittable"
)

// CameraParams is the derived state of (*camera.Camera).initialize
// (camera.go:179-253).  Package camera fills it after initialize() so that the
// C side never re-derives it in another precision.
type CameraParams struct {
	Width, Height, SppSqrt, MaxDepth               int
	Center, Pixel00, DeltaU, DeltaV, DefU, DefV, Bg [3]float64
	DefocusAngle, MaxContribution                   float64
}

// Variant names the render kernels (flag -variant): Auto lets the library pick per scene (the megakernel for list-only
// scenes such as the Cornell boxes, the wavefront kernels for scenes with a BVH), like the C++ and Python mirrors do.
type Variant int

const (
	Auto       Variant = C.GRT_VARIANT_AUTO
	Megakernel Variant = C.GRT_VARIANT_MEGAKERNEL
	Wavefront  Variant = C.GRT_VARIANT_WAVEFRONT
)

// Options selects devices and the kernel variant (flags -gpus, -variant).  The zero value renders on one GPU with
// the variant chosen by the library.
type Options struct {
	Seed    uint64
	Gpus    int
	Variant *Variant // nil: Auto
}

// lastError must run on the OS thread that made the failing call: the library keeps its error text per thread.
// Every exported function of this package therefore pins its goroutine with runtime.LockOSThread for the duration
// of its C calls (a goroutine that migrated between the call and grt_last_error would read an empty message).
func lastError(rc C.int) error {
	return fmt.Errorf("libgrt_cuda error %d: %s", int(rc), C.GoString(C.grt_last_error()))
}

// bvhOrder binds grt_bvh_order: BuildBVH's recursive sorts on the device (bvh.go:35-61).
func bvhOrder(boxes []float64) ([]uint32, error) {
	runtime.LockOSThread()
	defer runtime.UnlockOSThread()
	n := len(boxes) / 6
	order := make([]uint32, n)
	if n == 0 {
		return order, nil
	}
	if rc := C.grt_bvh_order((*C.double)(unsafe.Pointer(&boxes[0])), C.uint32_t(n), 0, (*C.uint32_t)(unsafe.Pointer(&order[0]))); rc != 0 {
		return nil, lastError(rc)
	}
	return order, nil
}

func init() {
	if DeviceCount() > 0 {
		hittable.BVHOrderHook = bvhOrder
	}
}

// DeviceCount reports the number of usable CUDA devices (0: the backend cannot be used; there is no CPU fallback in it).
func DeviceCount() int { return int(C.grt_device_count()) }

// Render replaces threadedRenderer/syncRenderer (camera.go:112-153): it renders
// the flattened scene and writes the P3 body ("%d %d %d\n" per pixel, color.go:45)
// to out.  tick is called H+1 times in total so that the progress bar finishes
// (progress.go:47-54, camera.go:107,131).
func Render(fs *hittable.FlatScene, cp CameraParams, opt Options, out io.Writer, tick func()) error {
	runtime.LockOSThread() // grt_last_error is per OS thread; the CUDA context of grt_render is too
	defer runtime.UnlockOSThread()
	if DeviceCount() == 0 {
		return errors.New("no CUDA device")
	}
	var sc C.GrtScene
	pin := fs.Fill(unsafe.Pointer(&sc)) // fills the POD view; pin keeps the Go slices alive and pinned
	defer pin.Unpin()

	var cam C.GrtCamera
	cam.width, cam.height = C.int32_t(cp.Width), C.int32_t(cp.Height)
	cam.spp_sqrt, cam.max_depth = C.int32_t(cp.SppSqrt), C.int32_t(cp.MaxDepth)
	for i := 0; i < 3; i++ {
		cam.center[i], cam.pixel00[i] = C.double(cp.Center[i]), C.double(cp.Pixel00[i])
		cam.delta_u[i], cam.delta_v[i] = C.double(cp.DeltaU[i]), C.double(cp.DeltaV[i])
		cam.defocus_u[i], cam.defocus_v[i] = C.double(cp.DefU[i]), C.double(cp.DefV[i])
		cam.background[i] = C.double(cp.Bg[i])
	}
	cam.defocus_angle, cam.max_contribution = C.double(cp.DefocusAngle), C.double(cp.MaxContribution)

	var o C.GrtOptions
	o.seed = C.uint64_t(opt.Seed)
	o.sample_stride = 1
	o.variant = C.GRT_VARIANT_AUTO
	if opt.Variant != nil {
		o.variant = C.int32_t(*opt.Variant)
	}

	n := cp.Width * cp.Height * 3
	sum := make([]float32, n)
	rgb := make([]byte, n)
	var rc C.int
	if opt.Gpus <= 1 {
		var h C.GrtSceneHandle
		if rc = C.grt_scene_upload(&sc, 0, &h); rc != 0 {
			return lastError(rc)
		}
		defer C.grt_scene_free(h)
		rc = C.grt_render(h, &cam, &o, (*C.float)(unsafe.Pointer(&sum[0])), (*C.uint8_t)(unsafe.Pointer(&rgb[0])), nil)
	} else {
		devs := make([]C.int, opt.Gpus)
		for i := range devs {
			devs[i] = C.int(i)
		}
		var ms C.double
		rc = C.grt_render_multi(&sc, &cam, &o, &devs[0], C.int(opt.Gpus),
			(*C.float)(unsafe.Pointer(&sum[0])), (*C.uint8_t)(unsafe.Pointer(&rgb[0])), &ms)
	}
	if rc != 0 {
		return lastError(rc)
	}
	// P3 body, row by row, exactly the text PrintColor writes (color.go:45)
	buf := make([]byte, 0, cp.Width*12)
	for row := 0; row < cp.Height; row++ {
		buf = buf[:0]
		for col := 0; col < cp.Width; col++ {
			p := rgb[(row*cp.Width+col)*3:]
			buf = fmt.Appendf(buf, "%d %d %d\n", p[0], p[1], p[2])
		}
		if _, err := out.Write(buf); err != nil {
			return err
		}
		tick()
	}
	tick()
	return nil
}
