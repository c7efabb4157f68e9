p='multigridbarriermpi.jl_b200/csrc/kernels.cuh'
s=open(p).read()
s=s.replace('''    const int32_t* lcols;    // [NU][E][LPE]
    const double* opd;       // [dim][B][nloc]
    const double* idd;       // coarse [NU][B][nloc]
    const double* own_val;   // fine [NU][nloc]
    const uint8_t* own_lq;   // fine [NU][nloc]
    const double* w;         // nloc''','''    const int32_t* lcols;    // [E][NU][LPE] element -> dof (-1: eliminated)
    const double* prec;      // [nloc][RW] per-point record: derivative rows (D*B), w, then
                             //   fine:   own_val[NU], own_lq bytes packed in one 8-byte slot
                             //   coarse: dense id-like rows [NU][B]
                             // RW even -> every record is 16-byte aligned (128-bit loads)''')
a=s.index("    // ---- gather the element's unknowns: lane q holds z[var][q]")
b=s.index("    // ---- apply_D: Dz = Dz0 + (D R) s")
new='''    // ---- loads.  Order matters (in-order issue): first the dof indices (their consumer, the gather
    // of s, comes last), then every independent stream, so one memory latency covers all of them.
    int32_t col[NU];
#pragma unroll
    for (int v = 0; v < NU; ++v) col[v] = act_e ? __ldg(&P.lcols[(e * NU + v) * LPE + l]) : -1;
    constexpr int RWF = D * B + 1 + NU + 1, RWC = D * B + 1 + NU * B;
    constexpr int RW = ((FINE ? RWF : RWC) + 1) / 2 * 2;
    double rec[RW];
    {
        const double2* __restrict__ rp = reinterpret_cast<const double2*>(P.prec + i * RW);
#pragma unroll
        for (int j = 0; j < RW / 2; ++j) {
            const double2 t2 = act ? __ldg(rp + j) : make_double2(0.0, 0.0);
            rec[2 * j] = t2.x;
            rec[2 * j + 1] = t2.y;
        }
    }
    double cc[ND], dz[ND];
#pragma unroll
    for (int k = 0; k < ND; ++k) {
        cc[k] = act ? __ldg(&P.c[(int64_t)k * n + i]) : 0.0;
        dz[k] = (act && P.Dz0) ? __ldg(&P.Dz0[(int64_t)k * n + i]) : 0.0;
    }
    // ---- gather the element's unknowns: lane q holds z[var][q]
    double zl[NU];
#pragma unroll
    for (int v = 0; v < NU; ++v) zl[v] = (col[v] >= 0) ? __ldg(&P.s[col[v]]) : 0.0;
    // ---- unpack the record
    double a[D][B];
#pragma unroll
    for (int k = 0; k < D; ++k)
#pragma unroll
        for (int q = 0; q < B; ++q) a[k][q] = rec[k * B + q];
    const double wi = rec[D * B];
    double aid[FINE ? 1 : NU][FINE ? 1 : B];
    double oval[NU];
    int olq[NU];
    bool oh[NU];  // this point owns a column of variable v (false: eliminated dof, e.g. Dirichlet)
    if (FINE) {
        const unsigned long long lqbits = (unsigned long long)__double_as_longlong(rec[FINE ? D * B + 1 + NU : 0]);
#pragma unroll
        for (int v = 0; v < NU; ++v) {
            oval[v] = rec[FINE ? D * B + 1 + v : 0];
            olq[v] = act ? (int)((lqbits >> (8 * v)) & 0xFFull) : 255;
            oh[v] = olq[v] != 255;
            if (!oh[v]) { oval[v] = 0.0; olq[v] = 0; }
        }
    } else {
#pragma unroll
        for (int v = 0; v < NU; ++v)
#pragma unroll
            for (int q = 0; q < B; ++q) aid[FINE ? 0 : v][FINE ? 0 : q] = rec[FINE ? 0 : D * B + 1 + v * B + q];
    }
'''
s=s[:a]+new+s[b:]
open(p,'w').write(s)

p='multigridbarriermpi.jl_b200/csrc/plan_host.h'
h=open(p).read()
h=h.replace('''    std::vector<int32_t> lcols;    // [NU][E][LPE]  global dof or -1
    std::vector<double> opd;       // dense derivative rows [dim][B][nloc]
    std::vector<double> idd;       // coarse: dense id-like rows [NU][B][nloc]
    std::vector<double> own_val;   // fine: [NU][nloc]
    std::vector<uint8_t> own_lq;   // fine: [NU][nloc]  local column or 255''','''    std::vector<int32_t> lcols;    // [E][NU][LPE]  global dof or -1
    int RW = 0;                    // doubles per point record (even)
    std::vector<double> prec;      // [nloc][RW]: derivative rows (dim*B), w, then fine: own_val[NU] + packed
                                   // own_lq bytes (255 = none); coarse: dense id-like rows [NU][B]''')
h=h.replace('''void build_element_plan(const std::vector<HostCSR>& D, const HostCSR& R, int64_t n_global,
                        const BarrierDesc& bar, ElementPlan& out, bool want_hessian = true);''','''void build_element_plan(const std::vector<HostCSR>& D, const HostCSR& R, int64_t n_global, const double* w_local,
                        const BarrierDesc& bar, ElementPlan& out, bool want_hessian = true);''')
open(p,'w').write(h)
p='multigridbarriermpi.jl_b200/csrc/plan_host.cpp'
c=open(p).read()
c=c.replace('''void build_element_plan(const std::vector<HostCSR>& D, const HostCSR& R, int64_t n_global,
                        const BarrierDesc& bar, ElementPlan& P, bool want_hessian) {''','''void build_element_plan(const std::vector<HostCSR>& D, const HostCSR& R, int64_t n_global, const double* w_local,
                        const BarrierDesc& bar, ElementPlan& P, bool want_hessian) {''')
c=c.replace("P.lcols[((size_t)v * E + e) * LPE + q] = tmp[q];","P.lcols[((size_t)e * nu + v) * LPE + q] = tmp[q];")
c=c.replace("const int32_t* lc = &P.lcols[((size_t)v * E + e) * LPE];","const int32_t* lc = &P.lcols[((size_t)e * nu + v) * LPE];")
c=c.replace("const int32_t ga = P.lcols[((size_t)(a1 / B) * E + e) * LPE + a1 % B];","const int32_t ga = P.lcols[((size_t)e * nu + a1 / B) * LPE + a1 % B];")
c=c.replace("const int32_t gb = P.lcols[((size_t)(a2 / B) * E + e) * LPE + a2 % B];","const int32_t gb = P.lcols[((size_t)e * nu + a2 / B) * LPE + a2 % B];")
c=c.replace("const int32_t a = P.lcols[((size_t)v * E + e) * LPE + q];","const int32_t a = P.lcols[((size_t)e * nu + v) * LPE + q];")
a=c.index("    // operator rows in element-local columns")
b=c.index("    // structural pattern of R'(sum_jk D_j' diag D_k)R")
c=c[:a]+'''    // per-point records in element-local columns (one contiguous, 16-byte aligned record per point)
    {
        const int rwf = dim * (int)B + 1 + nu + 1, rwc = dim * (int)B + 1 + nu * (int)B;
        const int RW = ((fine ? rwf : rwc) + 1) / 2 * 2;
        P.RW = RW;
        P.prec.assign((size_t)nloc * RW, 0.0);
        for (int kd = 0; kd < dim; ++kd) {
            const HostCSR& A = Ek[1 + kd];
            for (int64_t i = 0; i < nloc; ++i)
                for (int64_t p = A.ptr[i]; p < A.ptr[i + 1]; ++p)
                    P.prec[(size_t)i * RW + kd * B + local_of(0, i / B, A.idx[p])] = A.val[p];
        }
        for (int64_t i = 0; i < nloc; ++i) P.prec[(size_t)i * RW + dim * B] = w_local[i];
        if (fine) {
            for (int64_t i = 0; i < nloc; ++i) {
                unsigned long long bits = 0;
                for (int v = 0; v < nu; ++v) {
                    const HostCSR& A = Ek[idop[v]];
                    unsigned lq = 255;
                    if (A.ptr[i + 1] > A.ptr[i]) {
                        P.prec[(size_t)i * RW + dim * B + 1 + v] = A.val[A.ptr[i]];
                        lq = (unsigned)local_of(v, i / B, A.idx[A.ptr[i]]);
                    }
                    bits |= (unsigned long long)lq << (8 * v);
                }
                double packed;
                std::memcpy(&packed, &bits, sizeof(double));
                P.prec[(size_t)i * RW + dim * B + 1 + nu] = packed;
            }
        } else {
            for (int v = 0; v < nu; ++v) {
                const HostCSR& A = Ek[idop[v]];
                for (int64_t i = 0; i < nloc; ++i)
                    for (int64_t p = A.ptr[i]; p < A.ptr[i + 1]; ++p)
                        P.prec[(size_t)i * RW + dim * B + 1 + v * B + local_of(v, i / B, A.idx[p])] = A.val[p];
            }
        }
    }

'''+c[b:]
open(p,'w').write(c)
p='multigridbarriermpi.jl_b200/csrc/mgb_b200.cu'
m=open(p).read()
m=m.replace("    DevBuf<double> d_opd, d_idd, d_ownval, d_w, d_sel, d_rel, d_part, d_scal_tmp;\n    DevBuf<uint8_t> d_ownlq;","    DevBuf<double> d_prec, d_w, d_sel, d_rel, d_part, d_scal_tmp;")
m=m.replace("""    P.lcols = pl->d_lcols.p; P.opd = pl->d_opd.p; P.idd = pl->d_idd.p;
    P.own_val = pl->d_ownval.p; P.own_lq = pl->d_ownlq.p; P.w = pl->d_w.p;""","""    P.lcols = pl->d_lcols.p; P.prec = pl->d_prec.p;""")
m=m.replace("mgb::build_element_plan(Dh, Rh, n, pl->bar, pl->ep, want_hess);","mgb::build_element_plan(Dh, Rh, n, w_host + row0, pl->bar, pl->ep, want_hess);")
m=m.replace("""            pl->d_lcols.upload(ep.lcols, st); pl->d_opd.upload(ep.opd, st);
            pl->d_idd.upload(ep.idd, st); pl->d_ownval.upload(ep.own_val, st); pl->d_ownlq.upload(ep.own_lq, st);""","""            pl->d_lcols.upload(ep.lcols, st); pl->d_prec.upload(ep.prec, st);""")
m=m.replace("pl->d_lcols.bytes() + pl->d_opd.bytes() + pl->d_idd.bytes() + pl->d_ownval.bytes() +\n                            pl->d_ownlq.bytes() +","pl->d_lcols.bytes() + pl->d_prec.bytes() +")
m=m.replace("            std::vector<double>().swap(ep.opd); std::vector<double>().swap(ep.idd);","            std::vector<double>().swap(ep.prec);")
open(p,'w').write(m)
