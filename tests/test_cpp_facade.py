"""The header-only C++ façade compiles against the C ABI and links with libminimcmc.so (no GPU needed to
build; the tiny program only calls the host-side init routine when no device is present)."""
import os
import subprocess
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_facade_compiles_links_and_runs(tmp_path):
    src = tmp_path / "t.cpp"
    src.write_text(textwrap.dedent("""
        #include <cstdio>
        #include <type_traits>
        #include "minimcmc.hpp"
        int main() {
            auto x = mmc::init_det(4, 2);
            std::printf("%.16g\\n", x[0]);
            try {
                std::vector<float> init(8, 0.5f);
                mmc::HMC h(mmc::target(MMC_T_ROSENBROCK_2D, 2, {1.0, 100.0}), init, 4, 2, 0.01, 5);
                auto s = h.set_seed(1).run(3, 1);
                std::printf("ran %zu\\n", s.data.size());
            } catch (const mmc::Error &e) {
                std::printf("err %d\\n", e.code);
            }
            return 0;
        }
    """))
    exe = tmp_path / "t"
    lib_dir = os.path.join(ROOT, "mini_mcmc_b200")
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                    "-L", lib_dir, "-l:libminimcmc.so", f"-Wl,-rpath,{lib_dir}"], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split("\n")
    assert abs(float(out[0]) - 0.8343975468437959) < 1e-15
    assert out[1].startswith("ran 24") or out[1] == "err -2"   # -2 = MMC_ERR_NO_DEVICE on a CPU-only box
