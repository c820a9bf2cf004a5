// Copy into the reference checkout as scheduler/gpu_parity_harness_test.go (see baseline/go/README.md of the
// evergreen_b200 repo).  Reads a fixture written by dump_fixture.py, runs the REAL planner / queue-info / allocator
// on it without MongoDB, times it over GOMAXPROCS goroutines and writes one JSON result line per distro.
package scheduler

import (
	"bufio"
	"context"
	"encoding/json"
	"os"
	"runtime"
	"sync"
	"testing"
	"time"

	"github.com/evergreen-ci/evergreen"
	"github.com/evergreen-ci/evergreen/model/distro"
	"github.com/evergreen-ci/evergreen/model/host"
	"github.com/evergreen-ci/evergreen/model/task"
	"github.com/evergreen-ci/evergreen/util"
	"github.com/stretchr/testify/require"
)

type fxTask struct {
	Id, Version, Project, BuildVariant, TaskGroup, Requester, ActivatedBy, DistroId string
	TaskGroupOrder, TaskGroupMaxHosts, NumDependents                                int
	Priority, ActivatedAgoNs, ScheduledAgoNs, ExpectedNs                            int64
	GenerateTask, OverrideDependencies                                              bool
	DependsOn                                                                       []string
}
type fxHost struct {
	Id   string
	Free bool
}
type fxDistro struct {
	Distro    string
	NowNs     int64
	Planner   map[string]float64
	Allocator map[string]float64
	Tasks     []fxTask
	Hosts     []fxHost
}
type fxResult struct {
	Distro     string
	Order      []string
	TotalValue []int64
	Info       interface{}
	NewHosts   int
	FreeHosts  int
	AllocErr   string
}

func buildTasks(fx *fxDistro, now time.Time) []task.Task {
	out := make([]task.Task, 0, len(fx.Tasks))
	for _, f := range fx.Tasks {
		t := task.Task{
			Id: f.Id, Version: f.Version, Project: f.Project, BuildVariant: f.BuildVariant, TaskGroup: f.TaskGroup,
			TaskGroupOrder: f.TaskGroupOrder, TaskGroupMaxHosts: f.TaskGroupMaxHosts, Priority: f.Priority, Requester: f.Requester,
			GenerateTask: f.GenerateTask, ActivatedBy: f.ActivatedBy, NumDependents: f.NumDependents, DistroId: f.DistroId,
			ActivatedTime: now.Add(-time.Duration(f.ActivatedAgoNs)), ScheduledTime: now.Add(-time.Duration(f.ScheduledAgoNs)),
			OverrideDependencies: f.OverrideDependencies, Status: evergreen.TaskUndispatched,
			ExpectedDuration: time.Duration(f.ExpectedNs),
			// a fresh prediction: FetchExpectedDuration returns it without touching the database
			DurationPrediction: util.CachedDurationValue{Value: time.Duration(f.ExpectedNs), TTL: 24 * time.Hour, CollectedAt: now},
		}
		for _, dep := range f.DependsOn {
			t.DependsOn = append(t.DependsOn, task.Dependency{TaskId: dep, Status: evergreen.TaskSucceeded})
		}
		out = append(out, t)
	}
	return out
}

func buildDistro(fx *fxDistro) *distro.Distro {
	p := fx.Planner
	gv := p["group_versions"] != 0
	d := &distro.Distro{Id: fx.Distro, Provider: evergreen.ProviderNameEc2Fleet}
	d.PlannerSettings = distro.PlannerSettings{
		Version: evergreen.PlannerVersionTunable, TargetTime: time.Duration(int64(p["target_time_ns"])), GroupVersions: &gv,
		PatchFactor: int64(p["patch_factor"]), PatchTimeInQueueFactor: int64(p["patch_time_in_queue_factor"]),
		CommitQueueFactor: int64(p["commit_queue_factor"]), MainlineTimeInQueueFactor: int64(p["mainline_time_in_queue_factor"]),
		ExpectedRuntimeFactor: int64(p["expected_runtime_factor"]), GenerateTaskFactor: int64(p["generate_task_factor"]),
		NumDependentsFactor: p["num_dependents_factor"], StepbackTaskFactor: int64(p["stepback_task_factor"]),
	}
	if p["includes_dependencies"] != 0 {
		d.DispatcherSettings.Version = evergreen.DispatcherVersionRevisedWithDependencies
	}
	if a := fx.Allocator; a != nil {
		d.HostAllocatorSettings = distro.HostAllocatorSettings{
			Version: evergreen.HostAllocatorUtilization, MinimumHosts: int(a["minimum_hosts"]), MaximumHosts: int(a["maximum_hosts"]),
			FutureHostFraction: a["future_host_fraction"],
		}
		if a["round_up"] != 0 {
			d.HostAllocatorSettings.RoundingRule = evergreen.HostAllocatorRoundUp
		}
		if a["waits_over_thresh_feedback"] != 0 {
			d.HostAllocatorSettings.FeedbackRule = evergreen.HostAllocatorWaitsOverThreshFeedback
		}
		d.Disabled = a["disabled"] != 0
	}
	return d
}

func TestGPUParityHarness(t *testing.T) {
	path := os.Getenv("EVG_FIXTURE")
	if path == "" {
		t.Skip("EVG_FIXTURE not set")
	}
	f, err := os.Open(path)
	require.NoError(t, err)
	defer f.Close()
	var fixtures []fxDistro
	sc := bufio.NewScanner(f)
	sc.Buffer(make([]byte, 1<<20), 1<<30)
	for sc.Scan() {
		var fx fxDistro
		require.NoError(t, json.Unmarshal(sc.Bytes(), &fx))
		fixtures = append(fixtures, fx)
	}
	require.NoError(t, sc.Err())

	ctx := context.Background()
	results := make([]fxResult, len(fixtures))
	nTasks := 0
	for i := range fixtures {
		nTasks += len(fixtures[i].Tasks)
	}
	work := make(chan int, len(fixtures))
	for i := range fixtures {
		work <- i
	}
	close(work)
	var wg sync.WaitGroup
	started := time.Now()
	for w := 0; w < runtime.GOMAXPROCS(0); w++ {
		wg.Add(1)
		go func() {
			defer wg.Done()
			for i := range work {
				fx := &fixtures[i]
				now := time.Now()
				d := buildDistro(fx)
				tasks := buildTasks(fx, now)
				plan := PrepareTasksForPlanning(ctx, d, tasks).Export(ctx)
				info := GetDistroQueueInfo(ctx, d.Id, plan, d.GetTargetTime(), TaskPlannerOptions{
					IncludesDependencies: d.DispatcherSettings.Version == evergreen.DispatcherVersionRevisedWithDependencies})
				r := fxResult{Distro: fx.Distro, Info: info}
				for _, pt := range plan {
					r.Order = append(r.Order, pt.Id)
					r.TotalValue = append(r.TotalValue, pt.SortingValueBreakdown.TotalValue)
				}
				if fx.Allocator != nil {
					hosts := make([]host.Host, 0, len(fx.Hosts))
					for _, h := range fx.Hosts {
						if h.Free { // idle hosts only: a running host makes the allocator call task.Find (MongoDB)
							hosts = append(hosts, host.Host{Id: h.Id, Distro: *d})
						}
					}
					n, free, aerr := UtilizationBasedHostAllocator(ctx, &HostAllocatorData{Distro: *d, ExistingHosts: hosts, DistroQueueInfo: info})
					r.NewHosts, r.FreeHosts = n, free
					if aerr != nil {
						r.AllocErr = aerr.Error()
					}
				}
				results[i] = r
			}
		}()
	}
	wg.Wait()
	elapsed := time.Since(started)
	t.Logf("reference Go path: %d distros, %d tasks in %s on GOMAXPROCS=%d -> %.3g tasks/s, %.3g distros/s",
		len(fixtures), nTasks, elapsed, runtime.GOMAXPROCS(0), float64(nTasks)/elapsed.Seconds(), float64(len(fixtures))/elapsed.Seconds())

	if out := os.Getenv("EVG_RESULTS"); out != "" {
		g, err := os.Create(out)
		require.NoError(t, err)
		defer g.Close()
		enc := json.NewEncoder(g)
		for i := range results {
			require.NoError(t, enc.Encode(&results[i]))
		}
	}
}
