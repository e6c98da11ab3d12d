// varstore_oracle.cpp -- TEST INFRASTRUCTURE ONLY (never linked or executed by the product).
//
// The exact libtorch calls behind tch 0.16's `VarStore::save` / `VarStore::load` (border-tch-agent/src/dqn/base.rs:348-362,
// sac/base.rs:313-345 call them for every `<model>.pt.tch`), restated from tch's torch_api.cpp:
//     at_save_multi:      torch::serialize::OutputArchive archive; archive.write(name, tensor, /*buffer=*/false) ...;
//                         archive.save_to(filename)
//     at_load_callback:   auto module = torch::jit::load(filename); for (p : module.named_parameters()) f(p.name, p.value)
// plus the real `torch::optim::Adam` step, so that the op-by-op restatement in oracle/agent_oracle.py (CppAdam) can be
// checked against the optimizer tch binds bit for bit.
//
//   varstore_oracle save <file> <name> <n> <d0..> ... : tensors filled with a deterministic pattern
//   varstore_oracle load <file>                        : prints "name ndim dims... sum first last" per parameter
//   varstore_oracle adam <n> <steps> <lr> <b1> <b2> <eps> <wd> <adamw> : reads p0[n] and grads[steps][n] (raw f32) from
//                         stdin, writes p_final[n], m[n], v[n] (raw f32) to stdout
#include <torch/script.h>
#include <torch/torch.h>
#include <cstdio>
#include <iostream>
#include <string>
#include <vector>

static torch::Tensor pattern(const std::vector<int64_t>& shape, int seed) {
    int64_t n = 1;
    for (auto d : shape) n *= d;
    auto t = torch::arange(n, torch::kFloat32).mul_(0.001f).add_((float)seed).sin_();
    return t.reshape(shape).contiguous();
}

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    std::string cmd = argv[1];
    if (cmd == "save") {
        torch::serialize::OutputArchive archive;
        int i = 3, seed = 1;
        while (i < argc) {
            std::string name = argv[i++];
            int nd = atoi(argv[i++]);
            std::vector<int64_t> shape;
            for (int k = 0; k < nd; ++k) shape.push_back(atoll(argv[i++]));
            archive.write(name, pattern(shape, seed++), /*is_buffer=*/false);
        }
        archive.save_to(argv[2]);
        return 0;
    }
    if (cmd == "load") {
        auto module = torch::jit::load(argv[2]);
        for (const auto& p : module.named_parameters()) {
            auto v = p.value.contiguous().to(torch::kFloat32);
            std::cout << p.name << " " << v.dim();
            for (auto d : v.sizes()) std::cout << " " << d;
            auto flat = v.flatten();
            printf(" %.9g %.9g %.9g\n", flat.sum().item<double>(), flat[0].item<double>(), flat[flat.numel() - 1].item<double>());
        }
        return 0;
    }
    if (cmd == "adam") {
        const int64_t n = atoll(argv[2]);
        const int steps = atoi(argv[3]);
        const double lr = atof(argv[4]), b1 = atof(argv[5]), b2 = atof(argv[6]), eps = atof(argv[7]), wd = atof(argv[8]);
        const bool adamw = atoi(argv[9]) != 0;
        std::vector<float> p0(n), g((size_t)steps * n);
        if (fread(p0.data(), 4, n, stdin) != (size_t)n) return 3;
        if (fread(g.data(), 4, g.size(), stdin) != g.size()) return 3;
        torch::set_num_threads(1);
        auto p = torch::from_blob(p0.data(), {n}, torch::kFloat32).clone().set_requires_grad(true);
        std::unique_ptr<torch::optim::Optimizer> opt;
        if (adamw) opt.reset(new torch::optim::AdamW({p}, torch::optim::AdamWOptions(lr).betas({b1, b2}).eps(eps).weight_decay(wd)));
        else opt.reset(new torch::optim::Adam({p}, torch::optim::AdamOptions(lr).betas({b1, b2}).eps(eps).weight_decay(wd)));
        for (int s = 0; s < steps; ++s) {
            opt->zero_grad();
            p.mutable_grad() = torch::from_blob(g.data() + (size_t)s * n, {n}, torch::kFloat32).clone();
            opt->step();
        }
        torch::Tensor m, v;
        auto& st = opt->state().at(p.unsafeGetTensorImpl());
        if (adamw) {
            auto& s = static_cast<torch::optim::AdamWParamState&>(*st);
            m = s.exp_avg(); v = s.exp_avg_sq();
        } else {
            auto& s = static_cast<torch::optim::AdamParamState&>(*st);
            m = s.exp_avg(); v = s.exp_avg_sq();
        }
        auto out = [&](const torch::Tensor& t) { auto c = t.detach().contiguous(); fwrite(c.data_ptr<float>(), 4, n, stdout); };
        out(p); out(m); out(v);
        return 0;
    }
    return 2;
}
