nvidia-smi -q | grep -i "persistence" | head -2
cat > /tmp/t.cu <<'EOC'
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
int main() {
  auto t0 = std::chrono::steady_clock::now();
  cudaFree(0);
  auto t1 = std::chrono::steady_clock::now();
  void *p; cudaMalloc(&p, 1 << 20);
  auto t2 = std::chrono::steady_clock::now();
  printf("cudaFree(0) %.3f s, first cudaMalloc %.3f s\n", std::chrono::duration<double>(t1 - t0).count(), std::chrono::duration<double>(t2 - t1).count());
  return 0;
}
EOC
nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/t /tmp/t.cu
for i in 1 2 3; do /tmp/t; done
python - <<'PY' &
import torch, time
torch.zeros(1).cuda(); torch.cuda.synchronize()
print("holder up", flush=True)
time.sleep(25)
PY
sleep 12
echo "with another process holding a context:"
for i in 1 2 3; do /tmp/t; done
wait
