"""Power-limited steady state of the C4 bf16 batched matmul: this backend's CTA-pair tcgen05 kernel against cuBLAS
(torch.matmul) on the SAME shape, each looped for `secs` seconds on a cooled-down GPU, with an NVML time series of SM
clock and power. Answers: is the gap between the burst and the sustained TFLOP/s the kernel's energy per flop, or the
thermal state the loop starts from?   usage: python tools/power_probe.py [secs] [order: ours,cublas | cublas,ours]"""
import ctypes
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class Series:
    def __init__(self):
        import pynvml as nv
        nv.nvmlInit()
        self.nv, self.h = nv, nv.nvmlDeviceGetHandleByIndex(0)
        self.rows, self._stop = [], threading.Event()

    def start(self):
        self.rows, self._stop = [], threading.Event()
        self.t0 = time.perf_counter()
        self.th = threading.Thread(target=self._loop, daemon=True)
        self.th.start()

    def _loop(self):
        nv, h = self.nv, self.h
        while not self._stop.is_set():
            self.rows.append((round(time.perf_counter() - self.t0, 3), nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM),
                              round(nv.nvmlDeviceGetPowerUsage(h) / 1000.0, 1),
                              nv.nvmlDeviceGetTemperature(h, nv.NVML_TEMPERATURE_GPU)))
            self._stop.wait(0.1)

    def stop(self):
        self._stop.set()
        self.th.join()
        return self.rows


def main():
    secs = float(sys.argv[1]) if len(sys.argv) > 1 else 4.0
    order = (sys.argv[2] if len(sys.argv) > 2 else "ours,cublas").split(",")
    import torch
    from axiom_b200 import capi, device as dev
    L = capi.lib()
    capi.check(L.axb_set_device(0))
    nb, m = 64, 4096
    flops = 2.0 * nb * m ** 3
    ser = Series()
    out = {}

    def ours():
        A = dev.cast(dev.randn((nb, m, m), 11), capi.BF16)
        B = dev.cast(dev.randn((nb, m, m), 12, 0.0, 1.0 / 64.0), capi.BF16)
        C = dev.DeviceArray.empty((nb, m, m), capi.BF16)
        ad, bd, cd = A.desc(), B.desc(), C.desc()
        by = ctypes.byref

        def call():
            capi.check(L.axb_matmul(by(ad), by(bd), 0, 0, None, by(cd), None))
        for _ in range(3):
            call()
        dev.synchronize()
        return call, dev.synchronize

    def cublas():
        a = torch.randn(nb, m, m, device="cuda", dtype=torch.bfloat16)
        b = (torch.randn(nb, m, m, device="cuda") / 64).to(torch.bfloat16)
        c = torch.empty(nb, m, m, device="cuda", dtype=torch.bfloat16)

        def call():
            torch.bmm(a, b, out=c)
        for _ in range(3):
            call()
        torch.cuda.synchronize()
        return call, torch.cuda.synchronize

    for name in order:
        time.sleep(8.0)  # cool down between the two loops
        call, sync = ours() if name == "ours" else cublas()
        time.sleep(2.0)
        ser.start()
        t0 = time.perf_counter()
        n = 0
        marks = []
        while time.perf_counter() - t0 < secs:
            for _ in range(20):
                call()
            sync()
            n += 20
            marks.append((round(time.perf_counter() - t0, 3), n))
        dt = time.perf_counter() - t0
        rows = ser.stop()
        # TFLOP/s per second of the loop
        per_s, last_t, last_n = [], 0.0, 0
        for t, k in marks:
            if t - last_t >= 0.5:
                per_s.append((t, round((k - last_n) * flops / (t - last_t) / 1e12, 1)))
                last_t, last_n = t, k
        out[name] = {"TFLOPs_mean": n * flops / dt / 1e12, "calls": n, "secs": dt, "TFLOPs_over_time": per_s,
                     "nvml_t_sm_mhz_power_w_temp_c": rows[::3]}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
