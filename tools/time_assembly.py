import sys, json, statistics
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tools')
import torch, numpy as np
import polydeal_b200 as pdl
from pd_workloads import CONFIGS, build_handler
import time
for name in sys.argv[1:]:
    shape="blocks"
    if name.endswith("m"): name,shape=name[:-1],"metis"
    cfg=CONFIGS[name]
    if shape=="metis":
        grid=pdl.Grid.hyper_cube(cfg["dim"],0.0,1.0,cfg["n"].bit_length()-1)
        groups=pdl.metis_agglomerates(grid,(cfg["n"]//cfg["b"])**cfg["dim"])
        ah=pdl.AgglomerationHandler(grid); ah.define_agglomerates(groups); ah.initialize_fe_values(cfg["nq"]); ah.distribute_agglomerated_dofs(pdl.FE_DGQ,cfg["p"])
    else:
        ah=build_handler(pdl,cfg,1)
    t_create=time.time()
    desc=ah.flatten(penalty_constant=-1.0 if cfg["C"] is None else cfg["C"])
    stream=torch.cuda.Stream(); torch.cuda.set_stream(stream)
    op=pdl.SIPOperator(desc,keepalive=ah); op.set_stream(stream.cuda_stream)
    t_create=time.time()-t_create
    tot=[]; kms=[]
    for s in range(6):
        a,e=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record(stream); op.assemble(stiffness=1.0,mass=cfg["mass"]); e.record(stream); e.synchronize()
        if s>=2: tot.append(a.elapsed_time(e)); kms.append(op.last_kernel_ms())
    N=op.m()
    x=torch.from_numpy(np.sin(0.37*np.arange(N))).cuda(); y=torch.empty_like(x)
    op.set_operator(pdl.ASSEMBLE_ALL,1.0,cfg["mass"])
    res={}
    for mode,key in ((pdl.VMULT_BLOCK_CSR,"csr_ms"),(pdl.VMULT_MATRIX_FREE,"mf_ms")):
        for _ in range(2): op.vmult_ptr(y.data_ptr(),x.data_ptr(),mode=mode)
        a,e=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(5): op.vmult_ptr(y.data_ptr(),x.data_ptr(),mode=mode)
        e.record(stream); e.synchronize(); res[key]=a.elapsed_time(e)/5
        res[key+"_sum"]=float(y.sum())
    print(json.dumps({"shape":shape,"create_s":t_create,"n_ifaces":int(desc.n_ifaces),"vmult":res,"config":name,"path":op.assembly_path,"assemble_ms":statistics.mean(tot),"dofs_per_s":op.m()/(statistics.mean(tot)*1e-3),"kernel_ms":kms[-1],"mem_GB":torch.cuda.max_memory_allocated()/1e9}))
    del op
