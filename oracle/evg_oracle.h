/*
 * evg_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A C++17 restatement of the evergreen-ci/evergreen scheduler hot path
 * (scheduler.PlanDistro: tunable planner + queue info + utilization host
 * allocator) on reference-shaped inputs: string ids, Go time semantics, map
 * keyed unit construction.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library, and
 * only as the checker / the timed CPU baseline.  The product (libevgsched.so)
 * never links or calls it.
 *
 * Parity status: PINNED for unit scoring, grouping, queue-info lengths and
 * allocator arithmetic by the reference's own known-answer tests (replayed
 * from tests/golden/ by tests/test_oracle_golden.py); UNPINNED (restatement
 * only; the reference cannot be built here: no Go toolchain, tests need
 * MongoDB) for large-scale ordering, tie order (arbitrary in the reference),
 * over-threshold / wait counters and IncludesDependencies=true.
 *
 * Every function cites the reference file:line it follows (paths relative to
 * the reference checkout root).
 */
#ifndef EVG_ORACLE_H
#define EVG_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Go zero time.Time (year 1) sentinel; 0 is the Unix epoch (non-zero for
 * time.Time.IsZero, zero for utility.IsZeroTime). All other times are ns
 * since the Unix epoch. */
#define EVO_TIME_ZERO INT64_MIN

/* Columnar string array: string i is buf[off[i] .. off[i+1]). */
typedef struct { const char* buf; const int64_t* off; } evo_strcol;

/* task.Task fields read on the path (model/task/task.go:83-350). */
typedef struct {
  int64_t n;
  evo_strcol id, version, project, build_variant, task_group, requester,
      activated_by, distro_id, status;
  const int64_t* priority;
  const int32_t* task_group_order;
  const int32_t* task_group_max_hosts;
  const int32_t* num_dependents;
  const uint8_t* generate_task;
  const uint8_t* override_dependencies;
  const uint8_t* blocked;             /* Task.Blocked() of this task (task.go:3649) */
  const int64_t* activated_time;
  const int64_t* ingest_time;
  const int64_t* scheduled_time;
  const int64_t* dependencies_met_time;
  const int64_t* expected_ns;         /* FetchExpectedDuration(ctx).Average (task.go:3519) */
  /* DependsOn, CSR over tasks */
  const int64_t* dep_off;             /* n+1 */
  evo_strcol dep_task_id;
  evo_strcol dep_status;              /* Dependency.Status requirement */
  /* result of the DB lookup for dependencies that are NOT in the queue */
  const uint8_t* dep_found;
  evo_strcol dep_task_status;
  const uint8_t* dep_task_blocked;
  /* Dependency.FinishedAt per dependency (model/task/task.go Dependency); NULL = all zero.  Read by
   * setDependenciesMetTime (task.go:673-684) when DependenciesMet evaluates the dependencies afresh. */
  const int64_t* dep_finished_at;
} evo_tasks;

/* distro.PlannerSettings (model/distro/distro.go:286-300) + the dispatcher bit */
typedef struct {
  int64_t patch_factor;
  int64_t patch_time_in_queue_factor;
  int64_t commit_queue_factor;
  int64_t mainline_time_in_queue_factor;
  int64_t expected_runtime_factor;
  int64_t generate_task_factor;
  int64_t stepback_task_factor;
  double  num_dependents_factor;
  int64_t target_time_ns;        /* PlannerSettings.TargetTime, 0 = unset */
  int32_t group_versions;        /* ShouldGroupVersions() */
  int32_t has_container_pool;    /* d.ContainerPool != "" */
  int32_t includes_dependencies; /* DispatcherSettings.Version == "revised-with-dependencies" */
  int32_t _pad;
} evo_planner_settings;

/* task.SortingValueBreakdown flattened (model/task/task.go:3990-4038) */
enum {
  EVO_BD_TASK_GROUP_LENGTH = 0, EVO_BD_TOTAL_VALUE,
  EVO_BD_P_INITIAL, EVO_BD_P_TASK_GROUP, EVO_BD_P_GENERATOR, EVO_BD_P_COMMIT_QUEUE,
  EVO_BD_R_COMMIT_QUEUE, EVO_BD_R_NUM_DEPENDENTS, EVO_BD_R_ESTIMATED_RUNTIME,
  EVO_BD_R_MAINLINE_WAIT, EVO_BD_R_STEPBACK, EVO_BD_R_PATCH, EVO_BD_R_PATCH_WAIT,
  EVO_BD_N
};

/* model.TaskGroupInfo (model/task_queue.go:22-47); name is reported as the
 * index of a task carrying that group string (or -1 for ""). */
typedef struct {
  int64_t name_task;  /* index into the plan's task table of a task with this group; -1 => "" */
  int64_t count, count_free, count_required, max_hosts;
  int64_t expected_duration;
  int64_t count_duration_over_threshold, count_wait_over_threshold;
  int64_t count_dep_filled_merge_queue_tasks;
  int64_t duration_over_threshold;
} evo_group_info;

/* model.DistroQueueInfo (model/task_queue.go:49-75) */
typedef struct {
  int64_t length, length_with_dependencies_met, count_dep_filled_merge_queue_tasks;
  int64_t expected_duration, max_duration_threshold;
  int64_t count_duration_over_threshold, duration_over_threshold, count_wait_over_threshold;
  int64_t secondary_queue;
  int64_t n_groups;
} evo_queue_info;

/* host.Host fields read on the path (model/host/host.go:79-88) + the result
 * of task.Find(ByIds(running tasks)) (utilization_based_host_allocator.go:337) */
typedef struct {
  int64_t n;
  evo_strcol running_task, running_task_group, running_task_bv,
      running_task_project, running_task_version;
  const int64_t* teardown_start_time;
  const uint8_t* rt_found;
  const int64_t* rt_expected_ns;
  const int64_t* rt_std_ns;
  const int64_t* rt_start_time;
} evo_hosts;

typedef struct {
  const char* provider;      /* distro.Provider */
  const char* rounding_rule; /* HostAllocatorSettings.RoundingRule */
  const char* feedback_rule; /* HostAllocatorSettings.FeedbackRule */
  int32_t disabled;
  int32_t minimum_hosts, maximum_hosts;
  int32_t has_pool, pool_max_containers, parent_found, parent_maximum_hosts;
  int32_t _pad;
  double future_host_fraction;
} evo_alloc_settings;

enum { EVO_OK = 0, EVO_ERR_FUTURE_FRACTION = 1, EVO_ERR_POOL_SIZE = 2, EVO_ERR_PARENT_MISSING = 3 };

/* ---- single pieces (for replaying the reference's unit tests) ---- */

/* Unit.info + unitInfo.value (scheduler/planner.go:209-337) over the given members. */
void evo_unit_value(const evo_tasks* t, const int64_t* members, int64_t n_members,
                    const evo_planner_settings* s, int64_t now, int64_t out_bd[EVO_BD_N]);

/* PrepareTasksForPlanning(...).Export (scheduler/planner.go:431-481), canonical
 * tie policy. Returns number of emitted tasks; out_order[r] = input index at
 * rank r; out_bd (may be NULL) = EVO_BD_N values per rank; *out_n_units = plan.Len(). */
int64_t evo_plan(const evo_tasks* t, const evo_planner_settings* s, int64_t now,
                 int64_t* out_order, int64_t* out_bd, int64_t* out_n_units);

/* Task.DependenciesMet against the in-queue cache (model/task/task.go:632-671). */
void evo_deps_met(const evo_tasks* t, uint8_t* out_met);

/* GetDistroQueueInfo (scheduler/scheduler.go:56-159) over tasks in `order`.
 * Returns number of groups written to out_groups (capacity n_order+1). */
int64_t evo_get_distro_queue_info(const evo_tasks* t, const int64_t* order, int64_t n_order,
                       const char* distro_id, int64_t threshold_ns, int32_t includes_dependencies,
                       int64_t now, evo_queue_info* out_info, evo_group_info* out_groups);

/* d.GetTargetTime() (model/distro/distro.go:422-440) */
int64_t evo_target_time(const evo_planner_settings* s);

/* calcNewHostsNeeded (utilization_based_host_allocator.go:268-296) */
int64_t evo_calc_new_hosts_needed(int64_t short_ns, int64_t threshold_ns, int64_t expected_free,
                                  int64_t n_long, int64_t n_overdue, int64_t n_merge_queue, int32_t round_down);

/* calcExistingFreeHosts (utilization_based_host_allocator.go:300-318) over all hosts. */
int32_t evo_calc_existing_free_hosts(const evo_hosts* h, double future_host_fraction,
                                     int64_t threshold_ns, int64_t now, int64_t* out_free);

/* UtilizationBasedHostAllocator (utilization_based_host_allocator.go:26-130).
 * group_names: names of `groups` (n_groups). Mutates groups[i].count_free /
 * count_required like the reference. Returns evo status. */
int32_t evo_allocate(const evo_hosts* h, const evo_alloc_settings* a, const evo_queue_info* info,
                     evo_group_info* groups, const evo_strcol* group_names, int64_t now,
                     int64_t* out_new_hosts, int64_t* out_free_hosts);

/* groupByTaskGroup (utilization_based_host_allocator.go:223-260): for each
 * host the bucket index: -1 => "", i>=0 => groups[i], <=-2 => a named group
 * without queue info (distinct names get distinct codes -2, -3, ...).
 * Returns number of buckets (distinct names over hosts and infos). */
int64_t evo_group_by_task_group(const evo_hosts* h, const evo_strcol* group_names, int64_t n_groups,
                                int64_t* out_host_bucket);

/* FetchExpectedDuration decision logic (model/task/task.go:3519-3590 +
 * util/cached_value.go:125-145); ttl 0 is replaced by the un-jittered 8 h. */
void evo_fetch_expected_duration(int64_t pred_value, int64_t pred_std, int64_t pred_ttl,
                                 int64_t pred_collected_at, int64_t expected_duration,
                                 int64_t expected_std, int64_t now, int32_t hist_found,
                                 int64_t hist_avg, int64_t hist_std,
                                 int64_t* out_avg, int64_t* out_std);

/* ---- whole job, many distros, std::thread pool (CPU baseline) ----
 * Distro d owns tasks [task_off[d], task_off[d+1]) and hosts [host_off[d], ..).
 * Runs plan -> queue info -> allocator per distro. Outputs: out_order (global
 * task slots, distro-local indices), out_total_value (per rank), out_info[d],
 * out_new/out_free/out_status[d], and (optionally) the TaskGroupInfos after the
 * allocator mutated them. */
void evo_job_batch(const evo_tasks* t, const int64_t* task_off, const evo_hosts* h,
                   const int64_t* host_off, const evo_planner_settings* ps,
                   const evo_alloc_settings* as, const char* const* distro_ids, int64_t n_distros,
                   int64_t now, int32_t n_threads, int32_t* out_order, int64_t* out_total_value,
                   evo_queue_info* out_info, int64_t* out_new, int64_t* out_free, int32_t* out_status,
                   evo_group_info* out_groups /* may be NULL; distro d's infos at slot task_off[d]+d.. */,
                   int64_t* out_bd /* may be NULL; EVO_BD_N values per ranked task */);

#ifdef __cplusplus
}
#endif
#endif
