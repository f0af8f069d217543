"""Stage-by-stage GPU parity check (debug tool; the pytest -m gpu suite is the gate).
Prints max-abs errors of every materialised intermediate against the CPU derivations."""
import os, sys, time, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from deepsolid_b200 import cell as C, network, hamiltonian, qmc
from oracle import deepsolid_oracle as O, forward_laplacian as FL

torch.set_num_threads(os.cpu_count() or 8)
dev = torch.device("cuda", 0)
quick = "--quick" in sys.argv


def err(a, b):
    a = a.detach().cpu(); b = b.detach().cpu()
    return float((a - b).abs().max()) if a.numel() else 0.0


def phase(name):
    print(f"\n=== {name} ===", flush=True)


def run(fn, *a):
    try:
        fn(*a)
    except Exception:
        traceback.print_exc()
        sys.stdout.flush()


def check_gemm():
    from deepsolid_b200 import _lib
    lib = _lib.load()
    for (m, n, k) in [(8, 2, 2), (130, 34, 18), (1000, 256, 320), (777, 432, 256), (4096, 256, 16)]:
        a = torch.randn(m, k, dtype=torch.float64, device=dev)
        b = torch.randn(k, n, dtype=torch.float64, device=dev)
        c = torch.zeros(m, n, dtype=torch.float64, device=dev)
        rc = lib.ds_dgemm_probe(0, a.data_ptr(), b.data_ptr(), c.data_ptr(), m, n, k, None)
        torch.cuda.synchronize()
        print(f"gemm {m}x{n}x{k} rc={rc} err={err(c, a @ b):.3e}", flush=True)
    if quick:
        return
    m, n, k = 148 * 128 * 4, 256, 320
    a = torch.randn(m, k, dtype=torch.float64, device=dev)
    b = torch.randn(k, n, dtype=torch.float64, device=dev)
    c = torch.zeros(m, n, dtype=torch.float64, device=dev)
    for name, f in [("ours", lambda: lib.ds_dgemm_probe(0, a.data_ptr(), b.data_ptr(), c.data_ptr(), m, n, k, None)),
                    ("cublas", lambda: torch.matmul(a, b, out=c))]:
        for _ in range(3): f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): f()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"dgemm {name}: {m}x{n}x{k} {ms:.3f} ms  {2.0*m*n*k/ms/1e9:.2f} TF/s", flush=True)
    a = torch.randn(8192, 8192, dtype=torch.float64, device=dev); b = torch.randn(8192, 8192, dtype=torch.float64, device=dev)
    c = torch.empty_like(a)
    for _ in range(2): torch.matmul(a, b, out=c)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): torch.matmul(a, b, out=c)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"cublas dgemm 8192^3: {ms:.2f} ms {2*8192**3/ms/1e9:.2f} TF/s", flush=True)


def check_system(name, B):
    sc = C.build_system(name)
    kl = C.make_klist(sc)
    Pn = O.init_params(np.random.default_rng(888), sc.original_cell.natm, sc.nelec)
    P = O.params_to_torch(Pn)
    X = torch.as_tensor(C.init_walkers(sc, B))
    nu, nd = sc.nelec; N = nu + nd; ND = 3 * N; NDp = (ND + 7) // 8 * 8
    A = sc.original_cell.natm; K0 = 4 * A + 8; H, Pp, D = 256, 32, 8; K1 = H + 2 * Pp
    net = network.make_solid_fermi_net(envelope_type="isotropic", full_det=False, klist=kl, simulation_cell=sc,
                                       determinants=8, method_name="eval_logdet")
    hp = net.apply.hotpath()
    Xd = X.to(dev)
    la_r, ang_r, ke_r, inter = FL.kinetic_forward_laplacian(P, X, sc, kl, want=True)
    # forward
    v = net.apply(P, Xd)
    torch.cuda.synchronize()
    print(f"[{name}] logabs err {err(v.real, la_r):.3e}  phase err {float(torch.angle(torch.exp(1j*(v.imag.cpu()-ang_r))).abs().max()):.3e}", flush=True)
    mats = hp.orbitals(Xd)
    for s in range(2):
        print(f"[{name}] orbitals spin{s} err {err(mats[s], inter[f'orb{s}']):.3e} (max |M| {float(inter[f'orb{s}'].abs().max()):.3e})", flush=True)
    # stage checks of the laplacian path
    el = hamiltonian.local_energy_seperate(net.apply, sc, mode="for")
    for stop in (0, 1):
        hp.debug_set("stop_layer", stop)
        try:
            el(P, Xd)
        except Exception as e:
            print("stop run raised", e)
        torch.cuda.synchronize()
        hv = hp.debug_buffer(f"V{stop}")[: B * N * K1].reshape(B, N, K1)[:, :, :H]
        hJ = hp.debug_buffer(f"J{stop}")[: B * N * NDp * K1].reshape(B, N, NDp, K1)[:, :, :ND, :H]
        hl = hp.debug_buffer(f"L{stop}")[: B * N * K1].reshape(B, N, K1)[:, :, :H]
        print(f"[{name}] layer{stop}: value {err(hv, inter[f'h{stop+1}_v']):.3e} jac {err(hJ, inter[f'h{stop+1}_J']):.3e} lap {err(hl, inter[f'h{stop+1}_l']):.3e} (max lap {float(inter[f'h{stop+1}_l'].abs().max()):.2e})", flush=True)
        if stop == 0:
            a0v = hp.debug_buffer("A0V")[: B * N * K0].reshape(B, N, K0)
            print(f"[{name}] ae feats {err(a0v[:, :, :4*A], inter['ae_v']):.3e}", flush=True)
    hp.debug_set("stop_layer", -1)
    ke, ew = el(P, Xd)
    torch.cuda.synchronize()
    hv = hp.debug_buffer("V0")[: B * N * K1].reshape(B, N, K1)[:, :, :H]
    hJ = hp.debug_buffer("J0")[: B * N * NDp * K1].reshape(B, N, NDp, K1)[:, :, :ND, :H]
    hl = hp.debug_buffer("L0")[: B * N * K1].reshape(B, N, K1)[:, :, :H]
    print(f"[{name}] layer2: value {err(hv, inter['h3_v']):.3e} jac {err(hJ, inter['h3_J']):.3e} lap {err(hl, inter['h3_l']):.3e}", flush=True)
    for s, ns in enumerate((nu, nd)):
        m = torch.view_as_complex(hp.debug_buffer(f"MAT{s}")[: B * D * ns * ns * 2].reshape(B, D, ns, ns, 2))
        lm = torch.view_as_complex(hp.debug_buffer(f"LAPM{s}")[: B * D * ns * ns * 2].reshape(B, D, ns, ns, 2))
        da = torch.view_as_complex(hp.debug_buffer(f"DA{s}")[: B * D * NDp * ns * ns * 2].reshape(B, D, NDp, ns, ns, 2))[:, :, :ND]
        print(f"[{name}] spin{s}: M {err(m, inter[f'orb{s}']):.3e} dM {err(da.permute(0,2,1,3,4), inter[f'dorb{s}']):.3e} lapM {err(lm, inter[f'lorb{s}']):.3e} (max {float(inter[f'lorb{s}'].abs().max()):.2e})", flush=True)
    print(f"[{name}] KINETIC err vs forward-laplacian {err(ke, ke_r):.3e}  |ke| max {float(ke_r.abs().max()):.3e}", flush=True)
    # oracle (autodiff) for a couple of walkers
    f = O.make_solid_fermi_net(kl, sc, method_name="eval_logdet")
    elo = O.local_energy_seperate(f, sc, mode="partition", partition_number=1)
    for b in range(min(B, 2)):
        k_o, e_o = elo(P, X[b])
        print(f"[{name}] walker {b}: oracle ke {complex(k_o):.10f} gpu {complex(ke[b].cpu()):.10f} |d|={abs(complex(k_o)-complex(ke[b].cpu())):.2e}; ewald oracle {float(e_o):.10f} gpu {float(ew[b]):.10f} |d|={abs(float(e_o)-float(ew[b])):.2e}", flush=True)
    # host-buffer path
    ke_h, ew_h = el(P, X)
    print(f"[{name}] host-path ke err {err(ke_h, ke):.3e}", flush=True)
    # mcmc with supplied noise
    steps = 3
    g = torch.Generator().manual_seed(5)
    xi = torch.randn(steps, B, 3 * N, generator=g, dtype=torch.float64)
    u = torch.rand(steps, B, generator=g, dtype=torch.float64)
    slog = network.make_solid_fermi_net(envelope_type="isotropic", full_det=False, klist=kl, simulation_cell=sc,
                                        determinants=8, method_name="eval_slogdet", hotpath=hp)
    step = qmc.make_mcmc_step(slog.apply, B, sc.lattice_vectors(), steps=steps)
    xn, pm, masks = step(P, Xd, (xi, u), 0.3, return_masks=True)
    fo = O.make_solid_fermi_net(kl, sc, method_name="eval_slogdet")
    ostep = O.make_mcmc_step(lambda p, xx: O.batch_apply(fo, p, xx), B, sc.lattice_vectors(), steps=steps)
    xo, pmo, mo = ostep(P, X, (xi, u), 0.3)
    print(f"[{name}] mcmc: masks equal {bool((masks.cpu().bool() == mo).all())} pmove {float(pm):.4f}/{float(pmo):.4f} x err {err(xn, xo):.3e}", flush=True)
    xn2, pm2 = step(P, Xd, 1234, 0.3)
    print(f"[{name}] mcmc philox pmove {float(pm2):.4f} launches {hp.launch_count()}", flush=True)


phase("gemm"); run(check_gemm)
for nm, B in ([("h4", 3)] if quick else [("h4", 3), ("lih_prim", 5), ("graphene8", 4), ("h10", 3)]):
    phase(nm); run(check_system, nm, B)
print("done", flush=True)
