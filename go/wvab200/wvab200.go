// Package wvab200 binds libwva_b200.so (include/wva_b200.h) into WVA's controller.
//
// NOT COMPILED IN THIS REPO'S CI: the build image has no Go toolchain.  It is the binding a
// WVA maintainer would add (see INTEGRATION.md); the same C ABI is exercised by the Python
// ctypes harness in tests/.
//
// It replaces, behind the existing adapter types,
//   - modelanalyzer.ModelAnalyzer.AnalyzeModel   (internal/modelanalyzer/analyzer.go:25-34)
//   - optimizer.VariantAutoscalingsEngine.Optimize (internal/optimizer/optimizer.go:30-54)
// i.e. Server.Calculate + Manager.Optimize + System.GenerateSolution of pkg/core, pkg/manager.
package wvab200

/*
#cgo CFLAGS: -I${SRCDIR}/../../include
#cgo LDFLAGS: -L${SRCDIR}/../../workload_variant_autoscaler_b200 -lwva_b200 -lcudart
#include <stdlib.h>
#include "wva_b200.h"
*/
import "C"

import (
	"context"
	"fmt"
	"runtime"
	"sync"
	"unsafe"

	llmdOptv1alpha1 "github.com/llm-d-incubation/workload-variant-autoscaler/api/v1alpha1"
	interfaces "github.com/llm-d-incubation/workload-variant-autoscaler/internal/interfaces"
	"github.com/llm-d-incubation/workload-variant-autoscaler/internal/utils"
	"github.com/llm-d-incubation/workload-variant-autoscaler/pkg/config"
)

const (
	accNone    = -1 // accelerator name ""
	accUnknown = -2 // a name that is not in the accelerator table
)

// Engine owns one wva_handle (one GPU). One solve at a time, like the reference L1.
type Engine struct {
	mu sync.Mutex
	h  *C.wva_handle
}

func NewEngine(device int) (*Engine, error) {
	var h *C.wva_handle
	if rc := C.wva_create(&h, C.int(device)); rc != C.WVA_OK {
		return nil, fmt.Errorf("wva_create: %s", C.GoString(C.wva_strerror(rc)))
	}
	e := &Engine{h: h}
	runtime.SetFinalizer(e, func(e *Engine) { e.Close() })
	return e, nil
}

func (e *Engine) Close() {
	e.mu.Lock()
	defer e.mu.Unlock()
	if e.h != nil {
		C.wva_destroy(e.h)
		e.h = nil
	}
}

// fleet is the flat SoA image of config.SystemSpec (System.SetFromSpec's joins, done with maps here).
type fleet struct {
	accNames, serverNames                                      []string
	accCost                                                    []float32
	accMult, accType, typeCap                                  []int32
	perfPresent                                                []uint8
	alpha, beta, gamma, delta                                  []float32
	accCount, maxBatch, atTokens                               []int32
	srvModel, srvPrio, srvMinRep, srvMaxBatch, inTok, outTok   []int32
	srvHasTarget, srvKeep                                      []uint8
	sloITL, sloTTFT, sloTPS, arrival, curCost                  []float32
	curAcc, curRep                                             []int32
	unlimited, delayedBestEffort                               bool
	saturationPolicy                                           int
}

func pack(spec *config.SystemSpec) *fleet {
	f := &fleet{}
	accIdx, typeIdx, modelIdx := map[string]int{}, map[string]int{}, map[string]int{}
	for _, a := range spec.Accelerators.Spec {
		if _, dup := accIdx[a.Name]; dup {
			continue
		}
		accIdx[a.Name] = len(f.accNames)
		f.accNames = append(f.accNames, a.Name)
		if _, ok := typeIdx[a.Type]; !ok {
			typeIdx[a.Type] = len(typeIdx)
		}
		f.accCost = append(f.accCost, a.Cost)
		f.accMult = append(f.accMult, int32(a.Multiplicity))
		f.accType = append(f.accType, int32(typeIdx[a.Type]))
	}
	f.typeCap = make([]int32, len(typeIdx))
	for _, c := range spec.Capacity.Count {
		if t, ok := typeIdx[c.Type]; ok {
			f.typeCap[t] = int32(c.Count)
		}
	}
	for _, pd := range spec.Models.PerfData {
		if _, ok := modelIdx[pd.Name]; !ok {
			modelIdx[pd.Name] = len(modelIdx)
		}
	}
	A, M := len(f.accNames), len(modelIdx)
	n := M * A
	f.perfPresent = make([]uint8, n)
	f.alpha, f.beta, f.gamma, f.delta = make([]float32, n), make([]float32, n), make([]float32, n), make([]float32, n)
	f.accCount, f.maxBatch, f.atTokens = make([]int32, n), make([]int32, n), make([]int32, n)
	for _, pd := range spec.Models.PerfData {
		a, ok := accIdx[pd.Acc]
		if !ok {
			continue
		}
		k := modelIdx[pd.Name]*A + a
		f.perfPresent[k] = 1
		f.alpha[k], f.beta[k] = pd.DecodeParms.Alpha, pd.DecodeParms.Beta
		f.gamma[k], f.delta[k] = pd.PrefillParms.Gamma, pd.PrefillParms.Delta
		f.accCount[k], f.maxBatch[k], f.atTokens[k] = int32(pd.AccCount), int32(pd.MaxBatchSize), int32(pd.AtTokens)
	}
	type class struct {
		prio    int
		targets map[string]config.ModelTarget
	}
	classes := map[string]class{}
	for _, sc := range spec.ServiceClasses.Spec {
		p := sc.Priority
		if p < config.DefaultHighPriority || p > config.DefaultLowPriority {
			p = config.DefaultServiceClassPriority
		}
		c := class{prio: p, targets: map[string]config.ModelTarget{}}
		for _, mt := range sc.ModelTargets {
			c.targets[mt.Model] = mt
		}
		classes[sc.Name] = c
	}
	for _, sv := range spec.Servers.Spec {
		f.serverNames = append(f.serverNames, sv.Name)
		m, ok := modelIdx[sv.Model]
		if !ok {
			m = -1
		}
		f.srvModel = append(f.srvModel, int32(m))
		cls := sv.Class
		if cls == "" {
			cls = config.DefaultServiceClassName
		}
		prio, has := config.DefaultServiceClassPriority, uint8(0)
		var t config.ModelTarget
		if c, ok := classes[cls]; ok {
			prio = c.prio
			if mt, ok := c.targets[sv.Model]; ok {
				t, has = mt, 1
			}
		}
		f.srvPrio = append(f.srvPrio, int32(prio))
		f.srvHasTarget = append(f.srvHasTarget, has)
		f.sloITL, f.sloTTFT, f.sloTPS = append(f.sloITL, t.SLO_ITL), append(f.sloTTFT, t.SLO_TTFT), append(f.sloTPS, t.SLO_TPS)
		keep := uint8(0)
		if sv.KeepAccelerator {
			keep = 1
		}
		f.srvKeep = append(f.srvKeep, keep)
		f.srvMinRep = append(f.srvMinRep, int32(sv.MinNumReplicas))
		f.srvMaxBatch = append(f.srvMaxBatch, int32(sv.MaxBatchSize))
		ld := sv.CurrentAlloc.Load
		f.arrival = append(f.arrival, ld.ArrivalRate)
		f.inTok, f.outTok = append(f.inTok, int32(ld.AvgInTokens)), append(f.outTok, int32(ld.AvgOutTokens))
		ca := int32(accNone)
		if name := sv.CurrentAlloc.Accelerator; name != "" {
			if a, ok := accIdx[name]; ok {
				ca = int32(a)
			} else {
				ca = accUnknown
			}
		}
		f.curAcc = append(f.curAcc, ca)
		f.curRep = append(f.curRep, int32(sv.CurrentAlloc.NumReplicas))
		f.curCost = append(f.curCost, sv.CurrentAlloc.Cost)
	}
	f.unlimited = spec.Optimizer.Spec.Unlimited
	f.delayedBestEffort = spec.Optimizer.Spec.DelayedBestEffort
	f.saturationPolicy = int(config.SaturatedAllocationPolicyEnum(spec.Optimizer.Spec.SaturationPolicy))
	return f
}

func f32p(s []float32) *C.float {
	if len(s) == 0 {
		return nil
	}
	return (*C.float)(unsafe.Pointer(&s[0]))
}
func i32p(s []int32) *C.int32_t {
	if len(s) == 0 {
		return nil
	}
	return (*C.int32_t)(unsafe.Pointer(&s[0]))
}
func u8p(s []uint8) *C.uint8_t {
	if len(s) == 0 {
		return nil
	}
	return (*C.uint8_t)(unsafe.Pointer(&s[0]))
}

// Solve = SetFromSpec + Server.Calculate for every server + Manager.Optimize + GenerateSolution.
func (e *Engine) Solve(spec *config.SystemSpec) (*config.AllocationSolution, error) {
	f := pack(spec)
	S := len(f.serverNames)
	var pin runtime.Pinner // the fleet struct holds Go pointers: pin them for the duration of the call
	defer pin.Unpin()
	var cf C.wva_fleet
	cf.n_acc, cf.n_types = C.int32_t(len(f.accNames)), C.int32_t(len(f.typeCap))
	cf.n_models, cf.n_servers = C.int32_t(len(f.perfPresent)/max(len(f.accNames), 1)), C.int32_t(S)
	cf.acc_cost, cf.acc_multiplicity, cf.acc_type, cf.type_capacity = f32p(f.accCost), i32p(f.accMult), i32p(f.accType), i32p(f.typeCap)
	cf.perf_present, cf.perf_alpha, cf.perf_beta, cf.perf_gamma, cf.perf_delta = u8p(f.perfPresent), f32p(f.alpha), f32p(f.beta), f32p(f.gamma), f32p(f.delta)
	cf.perf_acc_count, cf.perf_max_batch, cf.perf_at_tokens = i32p(f.accCount), i32p(f.maxBatch), i32p(f.atTokens)
	cf.srv_model, cf.srv_priority, cf.srv_has_target = i32p(f.srvModel), i32p(f.srvPrio), u8p(f.srvHasTarget)
	cf.srv_slo_itl, cf.srv_slo_ttft, cf.srv_slo_tps = f32p(f.sloITL), f32p(f.sloTTFT), f32p(f.sloTPS)
	cf.srv_keep_acc, cf.srv_min_replicas, cf.srv_max_batch = u8p(f.srvKeep), i32p(f.srvMinRep), i32p(f.srvMaxBatch)
	cf.srv_arrival_rpm, cf.srv_in_tokens, cf.srv_out_tokens = f32p(f.arrival), i32p(f.inTok), i32p(f.outTok)
	cf.srv_cur_acc, cf.srv_cur_replicas, cf.srv_cur_cost = i32p(f.curAcc), i32p(f.curRep), f32p(f.curCost)
	for _, p := range []unsafe.Pointer{unsafe.Pointer(cf.acc_cost), unsafe.Pointer(cf.acc_multiplicity), unsafe.Pointer(cf.acc_type),
		unsafe.Pointer(cf.type_capacity), unsafe.Pointer(cf.perf_present), unsafe.Pointer(cf.perf_alpha), unsafe.Pointer(cf.perf_beta),
		unsafe.Pointer(cf.perf_gamma), unsafe.Pointer(cf.perf_delta), unsafe.Pointer(cf.perf_acc_count), unsafe.Pointer(cf.perf_max_batch),
		unsafe.Pointer(cf.perf_at_tokens), unsafe.Pointer(cf.srv_model), unsafe.Pointer(cf.srv_priority), unsafe.Pointer(cf.srv_has_target),
		unsafe.Pointer(cf.srv_slo_itl), unsafe.Pointer(cf.srv_slo_ttft), unsafe.Pointer(cf.srv_slo_tps), unsafe.Pointer(cf.srv_keep_acc),
		unsafe.Pointer(cf.srv_min_replicas), unsafe.Pointer(cf.srv_max_batch), unsafe.Pointer(cf.srv_arrival_rpm), unsafe.Pointer(cf.srv_in_tokens),
		unsafe.Pointer(cf.srv_out_tokens), unsafe.Pointer(cf.srv_cur_acc), unsafe.Pointer(cf.srv_cur_replicas), unsafe.Pointer(cf.srv_cur_cost)} {
		if p != nil {
			pin.Pin(p)
		}
	}
	if f.unlimited {
		cf.unlimited = 1
	}
	if f.delayedBestEffort {
		cf.delayed_best_effort = 1
	}
	cf.saturation_policy = C.int32_t(f.saturationPolicy)
	C.wva_tunables_default(&cf.tun)
	cf.tun.max_queue_to_batch_ratio = C.int32_t(config.MaxQueueToBatchRatio)
	cf.tun.accel_penalty_factor = C.float(config.AccelPenaltyFactor)

	feasible := make([]uint8, S)
	acc, replicas, batch := make([]int32, S), make([]int32, S), make([]int32, S)
	cost, itl, ttft := make([]float32, S), make([]float32, S), make([]float32, S)
	var win C.wva_allocs
	win.feasible, win.acc, win.replicas, win.batch = u8p(feasible), i32p(acc), i32p(replicas), i32p(batch)
	win.cost, win.itl, win.ttft = f32p(cost), f32p(itl), f32p(ttft)
	for _, p := range []unsafe.Pointer{unsafe.Pointer(win.feasible), unsafe.Pointer(win.acc), unsafe.Pointer(win.replicas),
		unsafe.Pointer(win.batch), unsafe.Pointer(win.cost), unsafe.Pointer(win.itl), unsafe.Pointer(win.ttft)} {
		if p != nil {
			pin.Pin(p)
		}
	}
	e.mu.Lock()
	rc := C.wva_solve(e.h, &cf, nil, &win)
	msg := C.GoString(C.wva_last_error(e.h))
	e.mu.Unlock()
	if rc != C.WVA_OK {
		return nil, fmt.Errorf("wva_solve: %s: %s", C.GoString(C.wva_strerror(rc)), msg)
	}
	sol := &config.AllocationSolution{Spec: map[string]config.AllocationData{}}
	for s := 0; s < S; s++ { // GenerateSolution: pkg/core/system.go:303-319
		if feasible[s] == 0 {
			continue
		}
		name := ""
		if acc[s] >= 0 {
			name = f.accNames[acc[s]]
		}
		sol.Spec[f.serverNames[s]] = config.AllocationData{
			Accelerator: name, NumReplicas: int(replicas[s]), MaxBatch: int(batch[s]), Cost: cost[s],
			ITLAverage: itl[s], TTFTAverage: ttft[s], Load: spec.Servers.Spec[s].CurrentAlloc.Load,
		}
	}
	return sol, nil
}

// VariantAutoscalingsEngine is the drop-in for internal/optimizer.VariantAutoscalingsEngine
// (same method set: interfaces.VariantAutoscalingsEngine, internal/interfaces/interfaces.go:10-17).
type VariantAutoscalingsEngine struct {
	engine *Engine
	spec   *config.SystemSpec
}

func NewVariantAutoscalingsEngine(engine *Engine, spec *config.SystemSpec) *VariantAutoscalingsEngine {
	return &VariantAutoscalingsEngine{engine: engine, spec: spec}
}

// Optimize mirrors internal/optimizer/optimizer.go:30-54 (the analysis argument is ignored there too).
func (v *VariantAutoscalingsEngine) Optimize(ctx context.Context, vaList llmdOptv1alpha1.VariantAutoscalingList,
	_ map[string]*interfaces.ModelAnalyzeResponse) (map[string]llmdOptv1alpha1.OptimizedAlloc, error) {
	sol, err := v.engine.Solve(v.spec)
	if err != nil {
		return nil, err
	}
	if len(sol.Spec) == 0 {
		return nil, fmt.Errorf("no feasible allocations found for all variants: ")
	}
	out := make(map[string]llmdOptv1alpha1.OptimizedAlloc)
	for _, va := range vaList.Items {
		if oa, err := utils.CreateOptimizedAlloc(va.Name, va.Namespace, sol); err == nil {
			out[va.Name] = *oa
		}
	}
	return out, nil
}
