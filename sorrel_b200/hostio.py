"""Host-facing stepping: actions arrive in pinned HOST memory, rewards go back to the host.

`HostStepPipeline.submit(host_actions)` is the end-to-end form of `Environment.take_turn` for a
caller whose policy lives on the host: H2D copy of the step's actions, the fused step launch
(observations stay in HBM, written straight into the replay ring -- they are the learner's input and
never cross PCIe), D2H copy of the step's rewards.  The three stages run on three CUDA streams with
`depth` buffers in flight, so the copies of step i+1 / i-1 overlap the kernel of step i.

With `graphs=True` the compute stage of every slot ([step kernel, device clock advance, index draw,
Buffer.sample gather]) is a captured CUDA graph over that slot's fixed device buffers: per step the
host then issues two copies, one graph launch and a handful of event calls -- no per-step argument
marshalling or allocation -- which is what keeps a 1 ms step (one shard of eight) from being host-bound.
"""
from __future__ import annotations

import os

import torch


def bind_to_gpu_numa_node(device_index: int) -> dict:
    """Pin the calling process to the CPUs of the NUMA node its GPU hangs off, so that the pinned staging buffers it
    allocates afterwards (first touch) and the threads that fill them are local to that GPU's PCIe root.  With one
    rank per GPU on a two-socket host this keeps the per-step host traffic of the eight ranks off the inter-socket
    link.  Best effort: returns what it did (or why not) and never raises."""
    try:
        import pynvml

        pynvml.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = device_index
        if visible:
            ids = visible.split(",")
            if device_index < len(ids) and ids[device_index].isdigit():
                phys = int(ids[device_index])
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(phys)).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        if len(bus.split(":")[0]) == 8:  # NVML pads the PCI domain to 8 hex digits, sysfs uses 4
            bus = bus[4:]
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:
            return {"bound": False, "why": "no NUMA information for the device"}
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return {"bound": False, "why": "node CPUs are outside the allowed set"}
        os.sched_setaffinity(0, cpus)
        return {"bound": True, "numa_node": node, "cpus": len(cpus)}
    except Exception as exc:  # noqa: BLE001
        return {"bound": False, "why": repr(exc)}


class HostStepPipeline:
    def __init__(self, engine, buffer, turns: int, depth: int = 2, rng: str = "philox", timing: bool = False,
                 graphs: bool = False, sample_batch: int = 0, action_dtype: torch.dtype = torch.int32):
        """`action_dtype`: torch.int32, or torch.uint8 -- action indices fit a byte (ActionSpec sizes <= 8), which
        quarters the host-to-device traffic of a host policy."""
        if action_dtype not in (torch.int32, torch.uint8):
            raise ValueError("action_dtype must be torch.int32 or torch.uint8")
        self.eng, self.buf, self.T, self.depth, self.rng = engine, buffer, int(turns), depth, rng
        self.timing, self.kernel_events = timing, []  # (start, end) CUDA events around every step launch
        self.B = int(sample_batch)
        dev, N, A = engine.device, engine.N, engine.A
        shape = (self.T, N, A)
        self.d_actions = [torch.zeros(shape, dtype=action_dtype, device=dev) for _ in range(depth)]
        self.d_rewards = [torch.empty(shape, dtype=torch.float32, device=dev) for _ in range(depth)]
        self.h_rewards = [torch.empty(shape, dtype=torch.float32).pin_memory() for _ in range(depth)]
        self.s_h2d, self.s_d2h = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        mk = lambda: [torch.cuda.Event() for _ in range(depth)]  # noqa: E731
        self.ev_h2d, self.ev_compute, self.ev_d2h = mk(), mk(), mk()
        self.i = 0
        self.h2d_bytes = self.d_actions[0].numel() * self.d_actions[0].element_size()
        self.d2h_bytes = self.d_rewards[0].numel() * 4
        self.sampled = None  # the last step's Buffer.sample batch (device tensors), when sample_batch > 0
        self.graphs = None
        if graphs:
            from .graphs import DeviceClock, StepGraph

            clock = DeviceClock(dev)
            self.graphs = [StepGraph(engine, buffer, self.T, rng=rng, sample_batch=self.B, actions=self.d_actions[k],
                                     reward_out=self.d_rewards[k], clock=clock) for k in range(depth)]

    def submit(self, h_actions: torch.Tensor) -> torch.Tensor:
        """Enqueue one step driven by `h_actions` (pinned int32 [T, N, A]).  Returns the pinned host
        reward buffer this step will fill (valid after `drain()` or once `ev_d2h[slot]` has completed)."""
        k = self.i % self.depth
        cur = torch.cuda.current_stream(self.eng.device)
        if self.i >= self.depth:
            self.ev_d2h[k].synchronize()  # the host buffer of slot k is about to be reused
        with torch.cuda.stream(self.s_h2d):
            self.s_h2d.wait_event(self.ev_compute[k])  # kernel that read d_actions[k] is done
            self.d_actions[k].copy_(h_actions, non_blocking=True)
            self.ev_h2d[k].record(self.s_h2d)
        cur.wait_event(self.ev_h2d[k])
        cur.wait_event(self.ev_d2h[k])  # previous read-back of d_rewards[k] is done
        if self.timing:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(cur)
        if self.graphs is not None:
            self.graphs[k].replay_step()
        else:
            self.eng.step(self.T, buffer=self.buf, actions=self.d_actions[k], reward_out=self.d_rewards[k],
                          rng=self.rng)
        if self.timing:
            e1.record(cur)
            self.kernel_events.append((e0, e1))
        self.ev_compute[k].record(cur)
        with torch.cuda.stream(self.s_d2h):
            self.s_d2h.wait_event(self.ev_compute[k])
            self.h_rewards[k].copy_(self.d_rewards[k], non_blocking=True)
            self.ev_d2h[k].record(self.s_d2h)
        if self.B:  # the learner's batch of this step (device-resident), after the rewards are on their way
            self.sampled = self.graphs[k].replay_sample() if self.graphs is not None else \
                self.buf.sample_device(self.B)
        self.i += 1
        return self.h_rewards[k]

    def drain(self) -> float:
        """Wait for everything in flight; returns a checksum of the last step's host rewards."""
        self.s_h2d.synchronize()
        self.s_d2h.synchronize()
        torch.cuda.current_stream(self.eng.device).synchronize()
        last = (self.i - 1) % self.depth if self.i else 0
        return float(self.h_rewards[last].sum())
