/*
 * solve_smoke.c — a compiled, non-Python caller of the C ABI (include/wva_b200.h): the closest available
 * stand-in for the cgo binding (no Go toolchain in this image).  Fills a wva_fleet by hand from the reference's
 * single-VariantAutoscaling fixture (deploy/examples/vllm-emulator/vllme-setup/vllme-variantautoscaling.yaml:26-37,
 * deploy/configmap-*.yaml: A100 cost 40, alpha 20.58 beta 0.41 gamma 5.2 delta 0.1, maxBatch 4, Premium SLO
 * itl 24 / ttft 500; production flags unlimited + keepAccelerator + minReplicas 1), calls wva_analyze,
 * wva_solve and wva_summarize, and prints the records as integers and float32 bit patterns.
 *
 *   usage: solve_smoke <arrival_rpm> <in_tokens> <out_tokens>
 *   exit:  0 ok, 3 no CUDA device (WVA_ERR_NO_DEVICE), 1 any other failure
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "wva_b200.h"

static uint32_t bits(float x) {
    uint32_t u;
    memcpy(&u, &x, 4);
    return u;
}

int main(int argc, char **argv) {
    if (argc < 4) {
        fprintf(stderr, "usage: %s <arrival_rpm> <in_tokens> <out_tokens>\n", argv[0]);
        return 1;
    }
    /* one accelerator, one type, one model, one server */
    float acc_cost[1] = {40.0f};
    int32_t acc_mult[1] = {1}, acc_type[1] = {0}, type_capacity[1] = {0};
    uint8_t perf_present[1] = {1};
    float alpha[1] = {20.58f}, beta[1] = {0.41f}, gamma_[1] = {5.2f}, delta[1] = {0.1f};
    int32_t acc_count[1] = {1}, max_batch[1] = {4}, at_tokens[1] = {0};
    int32_t srv_model[1] = {0}, srv_priority[1] = {1};
    uint8_t has_target[1] = {1}, keep_acc[1] = {1};
    float slo_itl[1] = {24.0f}, slo_ttft[1] = {500.0f}, slo_tps[1] = {0.0f};
    int32_t min_replicas[1] = {1}, srv_max_batch[1] = {4};
    float arrival[1] = {(float)atof(argv[1])};
    int32_t in_tok[1] = {atoi(argv[2])}, out_tok[1] = {atoi(argv[3])};
    int32_t cur_acc[1] = {0}, cur_replicas[1] = {1};
    float cur_cost[1] = {40.0f};

    wva_fleet f;
    memset(&f, 0, sizeof f);
    f.n_acc = 1; f.acc_cost = acc_cost; f.acc_multiplicity = acc_mult; f.acc_type = acc_type;
    f.n_types = 1; f.type_capacity = type_capacity;
    f.n_models = 1; f.perf_present = perf_present; f.perf_alpha = alpha; f.perf_beta = beta;
    f.perf_gamma = gamma_; f.perf_delta = delta; f.perf_acc_count = acc_count; f.perf_max_batch = max_batch;
    f.perf_at_tokens = at_tokens;
    f.n_servers = 1; f.srv_model = srv_model; f.srv_priority = srv_priority; f.srv_has_target = has_target;
    f.srv_slo_itl = slo_itl; f.srv_slo_ttft = slo_ttft; f.srv_slo_tps = slo_tps; f.srv_keep_acc = keep_acc;
    f.srv_min_replicas = min_replicas; f.srv_max_batch = srv_max_batch; f.srv_arrival_rpm = arrival;
    f.srv_in_tokens = in_tok; f.srv_out_tokens = out_tok; f.srv_cur_acc = cur_acc;
    f.srv_cur_replicas = cur_replicas; f.srv_cur_cost = cur_cost;
    f.unlimited = 1;
    wva_tunables_default(&f.tun);

    if (wva_abi_version() != WVA_ABI_VERSION) {
        fprintf(stderr, "header / library ABI mismatch: %d vs %d\n", WVA_ABI_VERSION, wva_abi_version());
        return 1;
    }
    wva_handle *h = NULL;
    int rc = wva_create(&h, 0);
    if (rc == WVA_ERR_NO_DEVICE) {
        printf("no-device %s\n", wva_strerror(rc));
        return 3;
    }
    if (rc != WVA_OK) {
        fprintf(stderr, "wva_create: %s\n", wva_strerror(rc));
        return 1;
    }

    uint8_t feas[2]; int32_t acc[2], rep[2], batch[2];
    float cost[2], value[2], itl[2], ttft[2], rho[2], max_rate[2];
    wva_allocs cand = {feas, acc, rep, batch, cost, value, itl, ttft, rho, max_rate};
    wva_allocs win = {feas + 1, acc + 1, rep + 1, batch + 1, cost + 1, value + 1, itl + 1, ttft + 1, rho + 1, max_rate + 1};

    rc = wva_analyze(h, &f, &cand); /* ModelAnalyzer.AnalyzeModel: Server.Calculate */
    if (rc != WVA_OK) {
        fprintf(stderr, "wva_analyze: %s: %s\n", wva_strerror(rc), wva_last_error(h));
        return 1;
    }
    printf("candidate %d %d %d %d %08x %08x %08x %08x %08x %08x\n", feas[0], acc[0], rep[0], batch[0], bits(cost[0]),
           bits(value[0]), bits(itl[0]), bits(ttft[0]), bits(rho[0]), bits(max_rate[0]));

    rc = wva_solve(h, &f, &cand, &win); /* VariantAutoscalingsEngine.Optimize */
    if (rc != WVA_OK) {
        fprintf(stderr, "wva_solve: %s: %s\n", wva_strerror(rc), wva_last_error(h));
        return 1;
    }
    printf("winner %d %d %d %d %08x %08x %08x %08x %08x %08x\n", feas[1], acc[1], rep[1], batch[1], bits(cost[1]),
           bits(value[1]), bits(itl[1]), bits(ttft[1]), bits(rho[1]), bits(max_rate[1]));

    uint8_t t_present[1]; int64_t t_count[1]; int32_t t_limit[1]; float t_cost[1];
    int32_t d_oa[1], d_na[1], d_or[1], d_nr[1]; float d_cost[1];
    wva_summary sum = {t_present, t_count, t_limit, t_cost, d_oa, d_na, d_or, d_nr, d_cost};
    rc = wva_summarize(h, &sum); /* System.AllocateByType + CreateAllocationDiff */
    if (rc != WVA_OK) {
        fprintf(stderr, "wva_summarize: %s: %s\n", wva_strerror(rc), wva_last_error(h));
        return 1;
    }
    printf("type %d %lld %d %08x\n", t_present[0], (long long)t_count[0], t_limit[0], bits(t_cost[0]));
    printf("diff %d %d %d %d %08x\n", d_oa[0], d_na[0], d_or[0], d_nr[0], bits(d_cost[0]));
    printf("launches %lld\n", (long long)wva_launch_count(h));
    wva_destroy(h);
    return 0;
}
