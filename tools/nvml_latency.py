import time, threading, torch, pynvml as n
n.nvmlInit(); h=n.nvmlDeviceGetHandleByIndex(0)
x=torch.randn(8192,8192,device='cuda'); stop=False
def work():
    while not stop:
        for _ in range(50): torch.mm(x,x)
        torch.cuda.synchronize()
t=threading.Thread(target=work); t.start()
calls={'clock':lambda: n.nvmlDeviceGetClockInfo(h,n.NVML_CLOCK_SM),'maxclock':lambda: n.nvmlDeviceGetMaxClockInfo(h,n.NVML_CLOCK_SM),'power':lambda: n.nvmlDeviceGetPowerUsage(h),'reasons':lambda: n.nvmlDeviceGetCurrentClocksEventReasons(h)}
for name,f in calls.items():
    lat=[]
    for i in range(60):
        t0=time.perf_counter(); f(); lat.append((time.perf_counter()-t0)*1e3); time.sleep(0.05)
    lat.sort(); print(name,'median %.2f ms p90 %.2f max %.2f'%(lat[len(lat)//2],lat[int(len(lat)*0.9)],lat[-1]))
stop=True; t.join()
