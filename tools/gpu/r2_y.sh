cat > /tmp/t.py <<'PY'
import sys, numpy as np
sys.path.insert(0, "/root/repo")
from axiom_b200 import capi, device as dev
from oracle import axiom_oracle as ao
capi.check(capi.lib().axb_set_device(0))
rng = np.random.default_rng(51)
H, W = int(sys.argv[1]), int(sys.argv[2])
x = rng.standard_normal((2, 64, H, W)).astype(np.float32); w = (rng.standard_normal((128, 64, 3, 3)) / 20).astype(np.float32)
dx = dev.DeviceArray.from_numpy(ao.f32_to_bf16(x), bf16=True); dw = dev.DeviceArray.from_numpy(ao.f32_to_bf16(w), bf16=True)
o = dev.conv(dx, dw, None, stride=(1,1), padding=(1,1), dilation=(1,1))
dev.synchronize()
print("ok", H, W, capi.lib().axb_matmul_last_path())
PY
for hw in "56 64" "56 128" "30 200" "56 56" "56 48"; do timeout 120 python /tmp/t.py $hw 2>&1 | tail -1 | cut -c1-200; done
