import pynvml
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
for name in ("NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_TX","NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_RX","NVML_FI_DEV_NVLINK_THROUGHPUT_RAW_TX","NVML_FI_DEV_NVLINK_LINK_COUNT"):
    fid = getattr(pynvml, name)
    for scope in (pynvml.NVML_NVLINK_MAX_LINKS, 0):
        try:
            v = pynvml.nvmlDeviceGetFieldValues(h, [(fid, scope)])[0]
            print(name, scope, "ret", v.nvmlReturn, "type", v.valueType, "ull", v.value.ullVal, "ui", v.value.uiVal)
        except Exception as e:
            print(name, scope, "EXC", repr(e))
try:
    for l in range(3):
        print("link", l, pynvml.nvmlDeviceGetNvLinkState(h, l))
except Exception as e:
    print("state EXC", repr(e))
