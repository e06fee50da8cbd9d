// Host side of the training backward: included by s3d.cu right after PlanBuilder.  build_backward() turns the tape a training plan
// recorded (PlanBuilder::BlockTape / UpTape) into the op list s3d_unet_backward replays — the adjoints of oracle/backward_ref.py in
// reverse forward order:
//     head -> decoder blocks (concat split, bilinear adjoint) -> encoder blocks (pool adjoint + skip joins) -> in_conv
// per block:  stage dOut -> [rollout adjoint, dgrad (tcgen05, k_conv_tc on flipped weights), wgrad] of conv2 -> GN+FiLM+SiLU backward
//             -> the same for conv1 -> GN+SiLU backward (+ identity skip) -> 1x1 skip dgrad.
// Gradient buffers live in a second arena (with reuse); parameter gradients accumulate in one flat fp32 buffer in state_dict order.

static void build_dgrad_packs(s3d_unet* u) {
    for (size_t i = 0; i < u->blocks.size(); ++i) {
        DevBlock& d = u->dblocks[i];
        for (DevConv3* cv : {&d.c1, &d.c2})
            for (int p = 0; p < 3; ++p) {
                const size_t n = static_cast<size_t>(cv->C) * 9 * cv->Cout;
                if (!cv->wd_pack[p]) cv->wd_pack[p] = dev_alloc<__half>(u->wallocs, 2 * n);
                launch_plain(k_pack_dgrad, dim3(static_cast<unsigned>(std::min<size_t>((n + 255) / 256, 1184))), dim3(256), 0, nullptr, cv->w_orig[p],
                             cv->Cout, cv->Cw, cv->C, cv->wd_pack[p]);
                LAUNCH_CHECK("k_pack_dgrad");
                for (int g = 1; g <= 2 && u->cfg.rollout; ++g) {
                    const size_t nv = static_cast<size_t>(9) * cv->Cout * cv->C;
                    if (!cv->wrv[p][g - 1]) cv->wrv[p][g - 1] = dev_alloc<float>(u->wallocs, nv);
                    launch_plain(k_pack_rollv, dim3(static_cast<unsigned>(std::min<size_t>((nv + 255) / 256, 1184))), dim3(256), 0, nullptr,
                                 cv->w_orig[p], cv->Cout, cv->C, g, roll_row_varying(p, g) ? 1 : 0, cv->wrv[p][g - 1]);
                    LAUNCH_CHECK("k_pack_rollv");
                }
                if (cv->Cs) {
                    const size_t ns = static_cast<size_t>(cv->Cs) * cv->Cout;
                    if (!cv->wsd_pack[p]) cv->wsd_pack[p] = dev_alloc<__half>(u->wallocs, 2 * ns);
                    launch_plain(k_pack_dgrad_1x1, dim3(static_cast<unsigned>((ns + 255) / 256)), dim3(256), 0, nullptr, cv->wskip_orig[p], cv->Cout,
                                 cv->Cs, cv->wsd_pack[p]);
                    LAUNCH_CHECK("k_pack_dgrad_1x1");
                }
            }
    }
    CUDA_TRY(cudaDeviceSynchronize());
}

static void build_grad_layout(s3d_unet* u) {
    u->grad_off.assign(u->tensors.size(), -1);
    int64_t n = 0;
    for (size_t i = 0; i < u->tensors.size(); ++i) {
        if (u->tensors[i].name == "__freqs") continue;
        u->grad_off[i] = n;
        n += (u->tensors[i].numel() + 3) / 4 * 4;
    }
    u->grad_numel = n;
}

struct BackwardBuilder {
    s3d_unet* u;
    PlanBuilder& pb;
    Plan* P;
    int B;

    void add(const char* name, double flops, std::function<void(cudaStream_t)> fn) {
        P->bwd_ops.push_back(std::move(fn));
        P->bwd_names.push_back(name);
        P->bwd_flops.push_back(flops);
    }
    float* grad(const std::string& name) {
        auto it = u->index.find(name);
        if (it == u->index.end()) throw S3dError{"internal: no gradient slot for " + name};
        return P->grads_own + u->grad_off[it->second];
    }
    void grad3(const std::string& prefix, const char* kind, const char* what, float* out[3]) {      // e.g. ("x.in_layers.2", "conv", "weight")
        for (int p = 0; p < 3; ++p) out[p] = grad(prefix + "." + kind + "_" + kPlane[p] + "." + what);
    }
    template <typename T>
    T* zeroed(size_t n) {       // reduction buffer cleared at the start of every backward
        T* p = dev_alloc<T>(P->allocs, n);
        P->bwd_zero.push_back({p, n * sizeof(T)});
        return p;
    }
    ActF allocF(int level, int C) {
        ActF a;
        a.C = C;
        a.level = level;
        for (int p = 0; p < 3; ++p) a.p.p[p] = static_cast<float*>(pb.barena.alloc(sizeof(float) * B * pb.px(level, p) * C, P->allocs));
        return a;
    }
    Act16 alloc16(int level, int C) {
        Act16 a;
        a.C = C;
        for (int p = 0; p < 3; ++p) a.p.p[p] = static_cast<__half*>(pb.barena.alloc(sizeof(__half) * 2 * B * pb.px(level, p) * C, P->allocs));
        return a;
    }
    void release(const ActF& a) {
        for (int p = 0; p < 3; ++p) pb.barena.release(a.p.p[p]);
    }
    void release(const Act16& a) {
        for (int p = 0; p < 3; ++p) pb.barena.release(a.p.p[p]);
    }
    // CTAs per (plane, sample) of the element-wise backward kernels: one wave at batch 1, about eight CTAs per SM in all at large
    // batch (their per-CTA reductions end in atomics on a few addresses: no more CTAs than fill the machine)
    int slots(int level) const {
        const int per_wave = std::min(2 * u->num_sms / 3, pb.max_px(level) / 32);
        return std::max(1, std::min(per_wave, std::max(4, 8 * u->num_sms / (3 * B))));
    }

    // ---- fp32 gradient of a conv output -> (hi, lo) pair + axis sums + per-channel totals
    struct Staged {
        Act16 pair;
        double* sums = nullptr;       // [B][total_len][C]
        double* bias_sum = nullptr;   // [3][C]
        double* bc_sum = nullptr;     // [B][C]
        PlanBuilder::Sums geo;        // segment offsets (shared with the forward's axis sums of the same level)
    };
    Staged stage(const ActF& g, int level, const PlanBuilder::Sums& geo, bool want_bc) {
        Staged S;
        S.pair = alloc16(level, g.C);
        S.geo = geo;
        S.sums = zeroed<double>(static_cast<size_t>(B) * geo.total_len * g.C);
        S.bias_sum = zeroed<double>(static_cast<size_t>(3) * g.C);
        if (want_bc) S.bc_sum = zeroed<double>(static_cast<size_t>(B) * g.C);
        StageArgs A{};
        A.g = PlanBuilder::cf(g.p);
        A.d = pb.dims[level];
        A.C = g.C;
        A.B = B;
        A.pair = S.pair.p;
        A.sums = S.sums;
        for (int i = 0; i < 6; ++i) A.seg_off[i] = geo.seg_off[i];
        A.total_len = geo.total_len;
        A.bias_sum = S.bias_sum;
        A.bc_sum = S.bc_sum;
        const int bx = g.C / 4, ny = std::max(1, 256 / bx);
        const TriDims d = pb.dims[level];
        const int strips = (std::max({d.rows[0], d.rows[1], d.rows[2]}) + kGsRows - 1) / kGsRows;
        const int Bv = B;
        add("k_grad_stage", 0.0, [=](cudaStream_t s) {
            launch_plain(k_grad_stage, dim3(strips, 3, Bv), dim3(bx, ny), sizeof(float) * ny * kGsRows * A.C, s, A);
            LAUNCH_CHECK("k_grad_stage");
        });
        return S;
    }

    // ---- backward of one 3x3 TriplaneConv (+ its fused 1x1 skip): rollout adjoint, dgrad, wgrad, bias
    // dY: fp32 gradient of the conv output; a: the conv's (hi, lo) input; fs: forward axis sums of a; returns dA (fp32, cv.C channels)
    ActF conv_backward(const ActF& dY, int level, const DevConv3& cv, const Act16& a, const PlanBuilder::Sums& fs, const Act16* x16,
                       const std::string& name, const std::string& skip_name, int emb_off, Staged* staged_out) {
        const bool ro = u->cfg.rollout;
        const TriDims d = pb.dims[level];
        Staged S = stage(dY, level, fs.buf ? fs : geo_of(level), emb_off >= 0);
        // bias gradients (and the additive-embedding gradient) straight from the staged totals
        {
            float *db[3], *dsb[3] = {nullptr, nullptr, nullptr};
            grad3(name, "conv", "bias", db);
            if (cv.Cs) grad3(skip_name, "conv", "bias", dsb);
            const int Cc = dY.C, Bv = B, fd = u->film_dim;
            const double* bs = S.bias_sum;
            const double* bc = S.bc_sum;
            float* dfilm = P->dfilm_own;
            add("k_bias_fin", 0.0, [=](cudaStream_t s) {
                launch_plain(k_bias_fin, dim3(1), dim3(256), 0, s, bs, Cc, db[0], db[1], db[2], bc, Bv, dfilm, fd, emb_off, dsb[0], dsb[1], dsb[2]);
                LAUNCH_CHECK("k_bias_fin");
            });
        }
        PlanBuilder::TBuf T{};
        if (ro) {
            for (int p = 0; p < 3; ++p) {
                T.Trow.p[p] = dev_alloc<float>(P->allocs, static_cast<size_t>(B) * 4 * d.rows[p] * cv.C);
                T.Tcol.p[p] = dev_alloc<float>(P->allocs, static_cast<size_t>(B) * 4 * d.cols[p] * cv.C);
            }
            RollBwdArgs R{};
            R.dy = PlanBuilder::cf(dY.p);
            R.d = d;
            R.sy = S.sums;
            R.fsums = fs.buf;
            R.total_len = fs.total_len;
            R.C = cv.C;
            R.Cout = cv.Cout;
            R.B = B;
            float* dw[3];
            grad3(name, "conv", "weight", dw);
            // (plane, group) -> source plane / which of its means (PlanBuilder::roll1d)
            static const int def[3][2][2] = {{{2, 0}, {1, 0}}, {{0, 0}, {2, 1}}, {{0, 1}, {1, 1}}};
            int Lmax = 0;
            for (int p = 0; p < 3; ++p)
                for (int g = 0; g < 2; ++g) {
                    RollBwdSrc& s = R.s[p * 2 + g];
                    const int sp = def[p][g][0], kind = def[p][g][1];
                    const bool rowv = roll_row_varying(p, g + 1);
                    s.plane = p;
                    s.g = g + 1;
                    s.row_varying = rowv ? 1 : 0;
                    s.L = rowv ? d.rows[p] : d.cols[p];
                    s.across_len = rowv ? d.cols[p] : d.rows[p];
                    s.sy_off = fs.seg_off[p * 2 + (rowv ? 0 : 1)];
                    s.vec_off = fs.seg_off[sp * 2 + kind];
                    const int avg_len = kind == 0 ? d.cols[sp] : d.rows[sp];
                    s.vec_scale = static_cast<float>(1.0 / 16777216.0 / static_cast<double>(avg_len));
                    s.inv_navg = static_cast<float>(1.0 / static_cast<double>(avg_len));
                    s.T = kind == 0 ? T.Trow.p[sp] : T.Tcol.p[sp];
                    s.wv = cv.wrv[p][g];
                    s.dw = dw[p];
                    S3D_CHECK((kind == 0 ? d.rows[sp] : d.cols[sp]) == s.L, "rollout adjoint geometry");
                    Lmax = std::max(Lmax, s.L);
                }
            const int Bv = B;
            S3D_CHECK(cv.C % 64 == 0 && cv.Cout % 64 == 0, "rollout adjoint tiling");
            const size_t smem = sizeof(float) * (3 * (kRvPos + 2) * cv.Cout + 32 * 64);
            S3D_CHECK(smem <= 200 * 1024, "k_roll_bwd_vec shared memory");
            const int nct = cv.C / 64;
            add("k_roll_bwd_vec", 2.0 * B * 6.0 * Lmax * cv.C * 9.0 * cv.Cout, [=](cudaStream_t s) {
                launch_plain(k_roll_bwd_vec, dim3((Lmax + kRvPos - 1) / kRvPos, 6 * nct, Bv), dim3(256), smem, s, R);
                LAUNCH_CHECK("k_roll_bwd_vec");
            });
            const int nsplit = std::min(B, 32);      // K = (sample, position) split by sample: 18 x nsplit CTAs, about four per SM at B = 32
            const size_t pn = static_cast<size_t>(nsplit) * 18 * 3 * cv.Cout * cv.C;
            float* partial = static_cast<float*>(pb.barena.alloc(sizeof(float) * pn, P->allocs));
            const int ntile = (cv.Cout / 64) * nct;
            add("k_roll_bwd_w", 2.0 * B * 6.0 * Lmax * cv.C * 9.0 * cv.Cout, [=](cudaStream_t s) {
                launch_plain(k_roll_bwd_w, dim3(ntile, 18, nsplit), dim3(256), 0, s, R, partial, nsplit);
                LAUNCH_CHECK("k_roll_bwd_w");
                launch_plain(k_roll_bwd_w_reduce, dim3((3 * cv.Cout * cv.C + 255) / 256, 18), dim3(256), 0, s, R, partial, nsplit);
                LAUNCH_CHECK("k_roll_bwd_w_reduce");
            });
            pb.barena.release(partial);
        }
        // dgrad: the forward kernel on flipped / transposed weights; the rollout adjoint rides in its epilogue as Trow / Tcol
        DevConv3 dg{};
        dg.C = cv.Cout;
        dg.Cout = cv.C;
        dg.Cw = cv.Cout;
        dg.Ktot = 9 * cv.Cout;
        for (int p = 0; p < 3; ++p) {
            dg.w_pack[p] = cv.wd_pack[p];
            dg.bias[p] = zero_bias(cv.C);
        }
        ActF dA = allocF(level, cv.C);
        {
            // route the launch into the backward list
            const size_t n0 = P->ops.size();
            pb.conv(S.pair, level, dg, ro ? &T : nullptr, nullptr, nullptr, -1, dA);
            move_last_op(n0, "k_conv_tc<dgrad>", pb.conv_flops(level, cv));
        }
        // wgrad (own channels) + skip 1x1 wgrad
        wgrad(S.pair, a, level, cv.C, cv.Cout, 9, cv.Cw, name, 2.0 * B * px3(level) * cv.Cout * 9.0 * cv.C);
        if (cv.Cs) wgrad(S.pair, *x16, level, cv.Cs, cv.Cout, 1, cv.Cs, skip_name, 2.0 * B * px3(level) * cv.Cout * cv.Cs);
        if (staged_out) *staged_out = S;
        else release(S.pair);
        return dA;
    }
    double px3(int level) const { return static_cast<double>(pb.px(level, 0) + pb.px(level, 1) + pb.px(level, 2)); }
    std::map<int, PlanBuilder::Sums> geo_cache;
    const PlanBuilder::Sums& geo_of(int level) {      // segment layout only (no device buffers): rollout-free models
        auto it = geo_cache.find(level);
        if (it != geo_cache.end()) return it->second;
        PlanBuilder::Sums S{};
        const TriDims d = pb.dims[level];
        int off = 0;
        for (int p = 0; p < 3; ++p) {
            S.seg_off[p * 2 + 0] = off;
            off += d.rows[p];
            S.seg_off[p * 2 + 1] = off;
            off += d.cols[p];
        }
        S.total_len = off;
        return geo_cache[level] = S;
    }
    std::map<int, float*> zero_bias_cache;
    float* zero_bias(int C) {
        auto it = zero_bias_cache.find(C);
        if (it != zero_bias_cache.end()) return it->second;
        float* p = dev_alloc<float>(P->allocs, C);
        CUDA_TRY(cudaMemset(p, 0, sizeof(float) * C));
        return zero_bias_cache[C] = p;
    }
    // PlanBuilder::conv appended its launch to the forward list: move it to the backward list
    void move_last_op(size_t n0, const char* name, double flops) {
        S3D_CHECK(P->ops.size() == n0 + 1, "internal: expected exactly one launch");
        P->bwd_ops.push_back(std::move(P->ops.back()));
        P->bwd_names.push_back(name);
        P->bwd_flops.push_back(flops);
        P->ops.pop_back();
        P->op_names.pop_back();
        P->op_flops.pop_back();
        P->op_trace.pop_back();
    }
    void wgrad(const Act16& dy, const Act16& a, int level, int C, int Cout, int ntap, int Cw, const std::string& name, double flops) {
        if (!u->bwd_wgrad_ffma && C % kBK == 0 && Cout % kBK == 0 && Cout <= 128) {
            wgrad_tc(dy, a, level, C, Cout, ntap, Cw, name, flops);
            return;
        }
        WgradArgs A{};
        A.dy = TriCH{{dy.p.p[0], dy.p.p[1], dy.p.p[2]}};
        A.a = TriCH{{a.p.p[0], a.p.p[1], a.p.p[2]}};
        A.d = pb.dims[level];
        A.C = C;
        A.Cout = Cout;
        A.B = B;
        A.ntap = ntap;
        A.Cw = Cw;
        grad3(name, "conv", "weight", A.dw);
        const long long total = static_cast<long long>(B) * pb.max_px(level);
        A.chunks = static_cast<int>(std::max<long long>(1, std::min<long long>(64, total / 256)));
        const int ny = ((Cout + 31) / 32) * ((C + 31) / 32) * ntap;
        add("k_wgrad_ffma", flops, [=](cudaStream_t s) {
            launch_plain(k_wgrad_ffma, dim3(A.chunks, ny, 3), dim3(32, 8), 0, s, A);
            LAUNCH_CHECK("k_wgrad_ffma");
        });
    }
    // tcgen05 version (wgrad_tc.cuh): MN-major operands straight from the NHWC pairs, split-K partials + fixed-order reduction
    void wgrad_tc(const Act16& dy, const Act16& a, int level, int C, int Cout, int ntap, int Cw, const std::string& name, double flops) {
        const TriDims d = pb.dims[level];
        auto maps = std::make_shared<WgradTcMaps>();
        memset(maps.get(), 0, sizeof(WgradTcMaps));
        WgradTcArgs A{};
        A.Cout = Cout;
        WgReduceArgs R{};
        const int n_acc = ntap == 9 ? 5 : 1;
        if (ntap == 9) {
            // tap pairs: (kh, kw 0|1) for kh = 0..2 (paired tap one pixel to the right), then kw = 2 of kh 0|1 (one row down) and
            // kw = 2 of kh = 2 with an unused second half
            const uint32_t W = kWgHaloW;
            const uint32_t off[5] = {0u, W * 128u, 2u * W * 128u, 2u * 128u, (2u * W + 2u) * 128u};
            const uint32_t lbo[5] = {128u, 128u, 128u, W * 128u, W * 128u};
            const int taps[5][2] = {{0, 1}, {3, 4}, {6, 7}, {2, 5}, {8, -1}};
            for (int i = 0; i < 5; ++i) {
                A.a_off[i] = off[i];
                A.lbo[i] = lbo[i];
                R.tap[i][0] = taps[i][0];
                R.tap[i][1] = taps[i][1];
            }
        } else {
            A.a_off[0] = (kWgHaloW + 1u) * 128u;       // centre of the halo patch
            A.lbo[0] = 128u;                      // second half unused
            R.tap[0][0] = 0;
            R.tap[0][1] = -1;
        }
        const int ngroups = (n_acc * Cout + 511) / 512;
        const int per_group = (n_acc + ngroups - 1) / ngroups;
        long long tiles[3];
        for (int p = 0; p < 3; ++p) {
            const uint64_t adims[5] = {static_cast<uint64_t>(C), static_cast<uint64_t>(d.cols[p]), static_cast<uint64_t>(d.rows[p]),
                                       static_cast<uint64_t>(B), 2};
            const uint32_t abox[5] = {kBK, kWgHaloW, kHaloH, 1, 1};
            make_tmap(&maps->a[p], a.p.p[p], 5, adims, abox);
            const uint64_t ydims[5] = {static_cast<uint64_t>(Cout), static_cast<uint64_t>(d.cols[p]), static_cast<uint64_t>(d.rows[p]),
                                       static_cast<uint64_t>(B), 2};
            const uint32_t ybox[5] = {kBK, kTileW, kTileH, 1, 1};
            make_tmap(&maps->y[p], dy.p.p[p], 5, ydims, ybox);
            A.tiles_x[p] = (d.cols[p] + kTileW - 1) / kTileW;
            A.tiles_per_sample[p] = A.tiles_x[p] * ((d.rows[p] + kTileH - 1) / kTileH);
            tiles[p] = static_cast<long long>(A.tiles_per_sample[p]) * B;
        }
        // split-K: about two waves of CTAs with equal MMA work
        const int cblks = C / kBK;
        double work_total = 0.0;
        for (int p = 0; p < 3; ++p) work_total += static_cast<double>(tiles[p]) * cblks * n_acc;
        const double quantum = work_total / (2.0 * u->num_sms);
        std::vector<WgUnit> units;
        std::vector<WgReduceGroup> groups;
        int max_nacc = 0;
        for (int p = 0; p < 3; ++p)
            for (int cb = 0; cb < cblks; ++cb)
                for (int g = 0; g < ngroups; ++g) {
                    const int acc0 = g * per_group, nacc = std::min(per_group, n_acc - acc0);
                    if (nacc <= 0) continue;
                    max_nacc = std::max(max_nacc, nacc);
                    long long ns = static_cast<long long>(static_cast<double>(tiles[p]) * nacc / quantum + 0.5);
                    ns = std::max<long long>(1, std::min<long long>(ns, tiles[p]));
                    groups.push_back(WgReduceGroup{static_cast<int>(units.size()), static_cast<int>(ns), p, cb, acc0, nacc});
                    for (long long s = 0; s < ns; ++s)
                        units.push_back(WgUnit{p, cb, acc0, nacc, static_cast<int>(tiles[p] * s / ns), static_cast<int>(tiles[p] * (s + 1) / ns)});
                }
        WgUnit* units_dev = dev_upload(P->allocs, units);
        WgReduceGroup* groups_dev = dev_upload(P->allocs, groups);
        const size_t partial_floats = units.size() * kWgMaxAcc * kBM * static_cast<size_t>(Cout);
        float* partial = static_cast<float*>(pb.barena.alloc(sizeof(float) * partial_floats, P->allocs));
        A.units = units_dev;
        A.partial = partial;
        R.groups = groups_dev;
        R.partial = partial;
        R.Cout = Cout;
        R.Cw = Cw;
        R.ntap = ntap;
        R.C = C;
        grad3(name, "conv", "weight", R.dw);
        const int n_units = static_cast<int>(units.size()), n_groups = static_cast<int>(groups.size());
        const size_t smem = 1024 + static_cast<size_t>(kWgStages) * (kWgABytes + (Cout / kBK) * kWgYBytes) + 128;
        S3D_CHECK(smem <= 227 * 1024, "k_wgrad_tc shared memory");
        add("k_wgrad_tc", flops, [=](cudaStream_t s) {
            launch_plain(k_wgrad_tc, dim3(n_units), dim3(kWgThreads), smem, s, *maps, A);
            LAUNCH_CHECK("k_wgrad_tc");
        });
        const int gx = (max_nacc * kBM * Cout + 255) / 256;
        add("k_wgrad_reduce", 0.0, [=](cudaStream_t s) {
            launch_plain(k_wgrad_reduce, dim3(gx, n_groups), dim3(256), 0, s, R);
            LAUNCH_CHECK("k_wgrad_reduce");
        });
        pb.barena.release(partial);
    }

    // ---- GroupNorm (+FiLM) + SiLU backward of one norm site
    ActF gn_backward(const ActF* x, const Act16* xh, int level, int C, const ActF& dy, const std::shared_ptr<SinkBox>& st, const ActF* add_g,
                     const std::string& norm_name) {
        ActF dx = allocF(level, C);
        GnBwdArgs A{};
        if (x) A.x = PlanBuilder::cf(x->p);
        else A.xh = TriCH{{xh->p.p[0], xh->p.p[1], xh->p.p[2]}};
        A.dy = PlanBuilder::cf(dy.p);
        A.d = pb.dims[level];
        A.C = C;
        A.B = B;
        A.acc = st->src.acc;
        A.gamma = st->src.gamma;
        A.beta = st->src.beta;
        A.film_dim = st->src.film_dim;
        A.film_off = st->src.film_off;
        A.psum = zeroed<double>(static_cast<size_t>(B) * 3 * C * 2);
        if (add_g) A.add = PlanBuilder::cf(add_g->p);
        A.dx = dx.p;
        A.nslots = slots(level);
        const bool use_film = st->use_film;
        Plan* Pp = P;
        const int bx = C / 4, ny = std::max(1, 256 / bx), Bv = B;
        add("k_gn_bwd_a", 0.0, [=](cudaStream_t s) {
            GnBwdArgs Al = A;
            if (use_film) {
                Al.film = Pp->film;
                Al.film_row = Pp->film_row;
            }
            launch_plain(k_gn_bwd_a, dim3(Al.nslots, 3, Bv), dim3(bx, ny), sizeof(float) * (4 + 2 * ny) * C, s, Al);
            LAUNCH_CHECK("k_gn_bwd_a");
        });
        add("k_gn_bwd_b", 0.0, [=](cudaStream_t s) {
            GnBwdArgs Al = A;
            if (use_film) {
                Al.film = Pp->film;
                Al.film_row = Pp->film_row;
            }
            launch_plain(k_gn_bwd_b, dim3(Al.nslots, 3, Bv), dim3(bx, ny), sizeof(float) * (6 * C + 64), s, Al);
            LAUNCH_CHECK("k_gn_bwd_b");
        });
        GnFinArgs F{};
        F.psum = A.psum;
        F.gamma = A.gamma;
        F.beta = A.beta;
        F.film_dim = A.film_dim;
        F.film_off = A.film_off;
        F.C = C;
        F.B = B;
        grad3(norm_name, "norm", "weight", F.dgamma);
        grad3(norm_name, "norm", "bias", F.dbeta);
        F.dfilm = P->dfilm_own;
        add("k_gn_bwd_fin", 0.0, [=](cudaStream_t s) {
            GnFinArgs Fl = F;
            if (use_film) {
                Fl.film = Pp->film;
                Fl.film_row = Pp->film_row;
            }
            launch_plain(k_gn_bwd_fin, dim3(std::max(1, (Fl.B * Fl.C + 255) / 256)), dim3(256), 0, s, Fl);
            LAUNCH_CHECK("k_gn_bwd_fin");
        });
        return dx;
    }

    // ---- one TriplaneResBlock; returns the gradient of the block input (fp32; of the 192-channel concat for a pair input)
    ActF block_backward(const PlanBuilder::BlockTape& T, const ActF& dOut) {
        const BlockSpec& b = u->blocks[T.bi];
        const DevBlock& w = u->dblocks[T.bi];
        const bool ssn = u->cfg.use_scale_shift_norm;
        const int level = T.level;
        Staged so{};
        ActF dA2 = conv_backward(dOut, level, w.c2, T.a2, T.s2, b.has_skip ? &T.x16 : nullptr, b.name + ".out_layers.2",
                                 b.name + ".skip_connection", -1, b.has_skip ? &so : nullptr);
        ActF dH1 = gn_backward(&T.h1, nullptr, level, b.cout, dA2, T.st2, nullptr, b.name + ".out_layers.0");
        release(dA2);
        ActF dA1 = conv_backward(dH1, level, w.c1, T.a1, T.s1, nullptr, b.name + ".in_layers.2", "", ssn ? -1 : b.film_off, nullptr);
        release(dH1);
        ActF dXm = gn_backward(T.x_is_pair ? nullptr : &T.x, T.x_is_pair ? &T.xh : nullptr, level, b.cin, dA1, T.st1,
                               b.has_skip ? nullptr : &dOut, b.name + ".in_layers.0");
        release(dA1);
        if (!b.has_skip) return dXm;
        // 1x1 skip dgrad: dX = dXm + Ws^T dOut as a pure 1x1 GEMM of the forward kernel (K = Cout chunks, residual = dXm)
        DevConv3 ds{};
        ds.C = 0;
        ds.Cs = b.cout;
        ds.Cout = b.cin;
        ds.Cw = 0;
        ds.Ktot = b.cout;
        for (int p = 0; p < 3; ++p) {
            ds.w_pack[p] = w.c2.wsd_pack[p];
            ds.bias[p] = zero_bias(b.cin);
        }
        ActF dX = allocF(level, b.cin);
        const size_t n0 = P->ops.size();
        pb.conv(so.pair, level, ds, nullptr, &so.pair, &dXm, -1, dX);
        move_last_op(n0, "k_conv_tc<dgrad 1x1>", 2.0 * B * px3(level) * b.cin * b.cout);
        release(so.pair);
        release(dXm);
        return dX;
    }
};

static void build_backward(s3d_unet* u, PlanBuilder& pb, const ActF& h_last, const std::shared_ptr<SinkBox>& st_head, const BoundaryArgs& bnd) {
    Plan* P = pb.P;
    const auto& c = u->cfg;
    const int B = pb.B, L = c.n_levels;
    S3D_CHECK(static_cast<int>(pb.enc_tape.size()) == L && static_cast<int>(pb.dec_tape.size()) == L, "internal: tape does not match the level structure");
    if (u->grad_numel == 0) build_grad_layout(u);
    P->grads_own = dev_alloc<float>(P->allocs, u->grad_numel);
    P->bwd_zero.push_back({P->grads_own, sizeof(float) * static_cast<size_t>(u->grad_numel)});
    P->dfilm_own = dev_alloc<float>(P->allocs, static_cast<size_t>(B) * u->film_dim);
    P->bwd_zero.push_back({P->dfilm_own, sizeof(float) * static_cast<size_t>(B) * u->film_dim});
    P->amax = dev_alloc<unsigned int>(P->allocs, 1);
    P->bwd_zero.push_back({P->amax, sizeof(unsigned int)});
    BackwardBuilder bb{u, pb, P, B};
    const int c0 = ch_of(c, 0), Cf = c.out_channels;
    const long long n_out = static_cast<long long>(B) * Cf * (bnd.H + bnd.Dd) * (bnd.W + bnd.Dd);
    const unsigned int* amax = P->amax;
    bb.add("k_grad_amax", 0.0, [=](cudaStream_t s) {
        launch_plain(k_grad_amax, dim3(static_cast<unsigned>(std::min<long long>((n_out + 255) / 256, 592))), dim3(256), 0, s, P->grad_out, n_out, P->amax);
        LAUNCH_CHECK("k_grad_amax");
    });
    // ---- head: dL/dout (composed) -> gradient of silu(GN(h)) + out.2 weight / bias gradients, then the out.0 norm backward
    ActF dy_head = bb.allocF(0, c0);
    {
        HeadBwdArgs A{};
        A.amax = amax;
        A.h = PlanBuilder::cf(h_last.p);
        A.d = pb.dims[0];
        A.C0 = c0;
        A.Cf = Cf;
        A.H = bnd.H; A.W = bnd.W; A.Dd = bnd.Dd;
        A.B = B;
        A.acc = st_head->src.acc;
        A.gamma = st_head->src.gamma;
        A.beta = st_head->src.beta;
        A.w_out = PlanBuilder::cf3(u->out_w);
        A.dy = dy_head.p;
        bb.grad3("out.2", "conv", "weight", A.dw);
        bb.grad3("out.2", "conv", "bias", A.db);
        for (int i = 0; i < 4; ++i) A.tile_start[i] = bnd.tile_start[i];
        for (int i = 0; i < 3; ++i) A.tiles_fast[i] = bnd.tiles_fast[i];
        S3D_CHECK((c0 == 64 || c0 == 128) && Cf <= kMaxCf, "head backward supports C0 = 64 | 128 and <= 16 triplane channels");
        bb.add("k_head_bwd", 4.0 * B * bb.px3(0) * c0 * Cf, [=](cudaStream_t s) {
            HeadBwdArgs Al = A;
            Al.g = P->grad_out;
            // two launches (dy, then dw / db): each half fits two CTAs per SM, the combined kernel held 236 registers
            const dim3 grid_dy(std::max(1, 8 * u->num_sms / (3 * B)), B, 3), grid_dw(std::max(1, 4 * u->num_sms / (3 * B)), B, 3);
            if (c0 == 64) {
                launch_plain(k_head_bwd<16, 0>, dim3(grid_dy), dim3(256), 0, s, Al);
                launch_plain(k_head_bwd<16, 1>, dim3(grid_dw), dim3(256), 0, s, Al);
            } else {
                launch_plain(k_head_bwd<32, 0>, dim3(grid_dy), dim3(256), 0, s, Al);
                launch_plain(k_head_bwd<32, 1>, dim3(grid_dw), dim3(256), 0, s, Al);
            }
            LAUNCH_CHECK("k_head_bwd");
        });
    }
    ActF dH = bb.gn_backward(&h_last, nullptr, 0, c0, dy_head, st_head, nullptr, "out.0");
    bb.release(dy_head);
    // ---- decoder, last block first
    std::vector<ActF> dskip(L);          // gradient reaching encoder level l's output through its decoder concat
    std::vector<bool> have_dskip(L, false);
    for (int j = L - 1; j >= 0; --j) {
        ActF dIn = bb.block_backward(pb.dec_tape[j], dH);
        bb.release(dH);
        if (j == 0) {
            dH = dIn;          // the deepest decoder block read the deepest encoder output directly
            break;
        }
        // dIn is the gradient of cat[resize(up2(low)), skip]
        const PlanBuilder::UpTape& U = pb.up_tape[j];
        ActF dlow = bb.allocF(U.low.level, U.low.C);
        dskip[U.out_level] = bb.allocF(U.out_level, U.skip.C);
        have_dskip[U.out_level] = true;
        UpcatBwdArgs A{};
        A.dcat = PlanBuilder::cf(dIn.p);
        A.dout = pb.dims[U.out_level];
        A.dlow = pb.dims[U.low.level];
        A.Cu = U.low.C;
        A.Cs = U.skip.C;
        A.B = B;
        A.do_up = U.do_up ? 1 : 0;
        A.dlow_g = dlow.p;
        A.dskip = dskip[U.out_level].p;
        A.nslots = bb.slots(U.out_level);
        const int Ct = A.Cu + A.Cs, bx = Ct / 4, ny = std::max(1, 256 / bx);
        size_t low_bytes[3];
        for (int p = 0; p < 3; ++p) low_bytes[p] = sizeof(float) * static_cast<size_t>(B) * pb.px(U.low.level, p) * U.low.C;
        // plain x2 (even sizes, no resize step): the up half as a deterministic gather over the low-resolution pixels
        bool plain_up = U.do_up;
        for (int p = 0; p < 3; ++p)
            plain_up = plain_up && 2 * A.dlow.rows[p] == A.dout.rows[p] && 2 * A.dlow.cols[p] == A.dout.cols[p];
        if (plain_up) {
            UpcatBwdArgs G = A;
            G.nslots = bb.slots(U.low.level);
            A.skip_only = 1;
            const int gbx = A.Cu / 4, gny = std::max(1, 256 / gbx);
            bb.add("k_up2_bwd_gather", 0.0, [=](cudaStream_t s) {
                launch_plain(k_up2_bwd_gather, dim3(G.nslots, 3, B), dim3(gbx, gny), 0, s, G);
                LAUNCH_CHECK("k_up2_bwd_gather");
            });
        }
        bb.add("k_upcat_bwd", 0.0, [=](cudaStream_t s) {
            if (!A.skip_only)
                for (int p = 0; p < 3; ++p) CUDA_TRY(cudaMemsetAsync(A.dlow_g.p[p], 0, low_bytes[p], s));      // scatter target
            launch_plain(k_upcat_bwd, dim3(A.nslots, 3, B), dim3(bx, ny), 0, s, A);
            LAUNCH_CHECK("k_upcat_bwd");
        });
        bb.release(dIn);
        dH = dlow;
    }
    // ---- encoder, deepest level first.  dH = gradient of level L-1's output.
    ActF dpool{};        // gradient of the pooled input of the level below (level l+1's block input)
    bool have_dpool = false;
    for (int l = L - 1; l >= 0; --l) {
        ActF dOut;
        if (l == L - 1) {
            dOut = dH;
        } else {
            // this level's output fed its decoder concat (dskip) and, through the 2x2 average pool, the next level (dpool)
            S3D_CHECK(have_dskip[l] && have_dpool, "internal: encoder gradient joins");
            dOut = bb.allocF(l, dskip[l].C);
            UpcatBwdArgs A{};           // Cu = 0: only the skip / pool-adjoint half of the kernel runs
            A.dcat = PlanBuilder::cf(dskip[l].p);
            A.dout = pb.dims[l];
            A.dlow = pb.dims[l];
            A.Cu = 0;
            A.Cs = dskip[l].C;
            A.B = B;
            A.dskip = dOut.p;
            A.dpool = PlanBuilder::cf(dpool.p);
            A.nslots = bb.slots(l);
            const int bx = A.Cs / 4, ny = std::max(1, 256 / bx);
            bb.add("k_upcat_bwd<pool adjoint>", 0.0, [=](cudaStream_t s) {
                launch_plain(k_upcat_bwd, dim3(A.nslots, 3, B), dim3(bx, ny), 0, s, A);
                LAUNCH_CHECK("k_upcat_bwd");
            });
            bb.release(dskip[l]);
            bb.release(dpool);
        }
        ActF dX = bb.block_backward(pb.enc_tape[l], dOut);
        bb.release(dOut);
        dpool = dX;
        have_dpool = true;
    }
    // ---- in_conv: weight / bias gradients only (x_t does not require a gradient in TrainLoop)
    {
        InconvBwdArgs A{};
        A.dh0 = PlanBuilder::cf(dpool.p);
        A.d = pb.dims[0];
        A.C0 = c0;
        A.Cf = c.in_channels;
        A.H = bnd.H; A.W = bnd.W; A.Dd = bnd.Dd;
        A.B = B;
        bb.grad3("in_conv.0", "conv", "weight", A.dw);
        bb.grad3("in_conv.0", "conv", "bias", A.db);
        for (int i = 0; i < 4; ++i) A.tile_start[i] = bnd.tile_start[i];
        for (int i = 0; i < 3; ++i) A.tiles_fast[i] = bnd.tiles_fast[i];
        bb.add("k_inconv_wgrad", 2.0 * B * bb.px3(0) * c0 * A.Cf, [=](cudaStream_t s) {
            InconvBwdArgs Al = A;
            Al.x = P->x;
            const dim3 grid(std::max(1, 4 * u->num_sms / (3 * B)), B, 3);
            if (c0 == 64) launch_plain(k_inconv_wgrad<16>, dim3(grid), dim3(256), 0, s, Al);
            else launch_plain(k_inconv_wgrad<32>, dim3(grid), dim3(256), 0, s, Al);
            LAUNCH_CHECK("k_inconv_wgrad");
        });
        bb.release(dpool);
    }
    // ---- un-scale
    {
        float* g = P->grads_own;
        float* df = P->dfilm_own;
        const long long n = u->grad_numel, nf = static_cast<long long>(B) * u->film_dim;
        bb.add("k_unscale", 0.0, [=](cudaStream_t s) {
            launch_plain(k_unscale, dim3(static_cast<unsigned>(std::min<long long>((n + 255) / 256, 1184))), dim3(256), 0, s, g, n, amax);
            launch_plain(k_unscale, dim3(static_cast<unsigned>(std::min<long long>((nf + 255) / 256, 64))), dim3(256), 0, s, df, nf, amax);
            LAUNCH_CHECK("k_unscale");
        });
    }
}
