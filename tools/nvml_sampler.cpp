// GPU utilisation side-car: what the reference's NVML/NVML.cpp (lines 18-92) does next to a training job —
// every ~1/6 s print, per GPU, a wall-clock stamp, the device name, SM utilisation, memory-controller utilisation and the
// bytes of memory in use; unbuffered stdout; SIGINT / SIGTERM end the loop. Written independently of the reference's
// source: NVML is resolved at run time (dlopen of libnvidia-ml.so.1), so the tool builds without the CUDA toolkit's nvml.h
// and exits with a clear message (code 2) on a box without a driver.
//   build: g++ -O2 -std=c++17 -o tools/nvml_sampler tools/nvml_sampler.cpp -ldl
//   use  : tools/nvml_sampler [--period-us 166667] [--count N] > /result/$MODEL/gpu.txt
#include <chrono>
#include <csignal>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <dlfcn.h>
#include <thread>

namespace {
volatile std::sig_atomic_t g_run = 1;
void on_signal(int) { g_run = 0; }

struct Utilization { unsigned int gpu, memory; };
struct Memory { unsigned long long total, free_, used; };
using Dev = void*;
template <typename F> bool load(void* lib, const char* name, F& fn) { fn = reinterpret_cast<F>(dlsym(lib, name)); return fn != nullptr; }
}  // namespace

int main(int argc, char** argv) {
  long period_us = 166667;  // the reference's period (NVML.cpp:84)
  long count = -1;
  for (int i = 1; i < argc; ++i) {
    if (!strcmp(argv[i], "--period-us") && i + 1 < argc) period_us = atol(argv[++i]);
    else if (!strcmp(argv[i], "--count") && i + 1 < argc) count = atol(argv[++i]);
    else { fprintf(stderr, "usage: %s [--period-us N] [--count N]\n", argv[0]); return 64; }
  }
  std::signal(SIGINT, on_signal);
  std::signal(SIGTERM, on_signal);
  setvbuf(stdout, nullptr, _IONBF, 0);
  void* lib = dlopen("libnvidia-ml.so.1", RTLD_NOW);
  if (!lib) { fprintf(stderr, "nvml_sampler: libnvidia-ml.so.1 not found (no NVIDIA driver on this box)\n"); return 2; }
  int (*init)() = nullptr; int (*shutdown)() = nullptr; int (*get_count)(unsigned int*) = nullptr;
  int (*get_handle)(unsigned int, Dev*) = nullptr; int (*get_name)(Dev, char*, unsigned int) = nullptr;
  int (*get_util)(Dev, Utilization*) = nullptr; int (*get_mem)(Dev, Memory*) = nullptr;
  if (!load(lib, "nvmlInit_v2", init) || !load(lib, "nvmlShutdown", shutdown) || !load(lib, "nvmlDeviceGetCount_v2", get_count) ||
      !load(lib, "nvmlDeviceGetHandleByIndex_v2", get_handle) || !load(lib, "nvmlDeviceGetName", get_name) ||
      !load(lib, "nvmlDeviceGetUtilizationRates", get_util) || !load(lib, "nvmlDeviceGetMemoryInfo", get_mem)) {
    fprintf(stderr, "nvml_sampler: NVML symbols missing\n");
    return 3;
  }
  if (init() != 0) { fprintf(stderr, "nvml_sampler: nvmlInit failed\n"); return 1; }
  unsigned int n = 0;
  if (get_count(&n) != 0) { shutdown(); return 1; }
  while (g_run && count != 0) {
    const auto t0 = std::chrono::steady_clock::now();
    const auto now = std::chrono::system_clock::now();
    const std::time_t tt = std::chrono::system_clock::to_time_t(now);
    std::tm tmv;
    localtime_r(&tt, &tmv);
    const int ms = (int)(std::chrono::duration_cast<std::chrono::milliseconds>(now.time_since_epoch()).count() % 1000);
    for (unsigned int i = 0; i < n; ++i) {
      Dev d;
      char name[96] = "?";
      Utilization u{0, 0};
      Memory m{0, 0, 0};
      if (get_handle(i, &d) != 0) continue;
      get_name(d, name, sizeof(name));
      const int ru = get_util(d, &u), rm = get_mem(d, &m);
      if (ru != 0 && rm != 0) continue;
      printf("%d:%d:%d:%d  Device %u: %s  GPU Util: %u  Mem Util: %u Mem Usage: %llu\n", tmv.tm_hour, tmv.tm_min, tmv.tm_sec, ms, i, name,
             u.gpu, u.memory, m.used);
    }
    if (count > 0) --count;
    const auto spent = std::chrono::steady_clock::now() - t0;
    const auto period = std::chrono::microseconds(period_us);
    if (spent < period) std::this_thread::sleep_for(period - spent);
  }
  shutdown();
  dlclose(lib);
  return 0;
}
