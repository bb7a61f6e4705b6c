// sgmof — command line of the reference's proj/ (proj/src/sgmof_main.cpp:11-78) on top of the facade:
//   sgmof I1 I2 [-o=flow.png] [-m=0|1] [-c=calib.txt] [-b=0|1] [-p=2] [-d] [-V] [-N=5] [-g=geometry.txt]
// Same positional arguments and key names; both "-k=value" (OpenCV's CommandLineParser form) and "-k value" are accepted.
// The reference stops after compute() ("//write optical flow", :75-77); here the flow is written as a KITTI 16-bit PNG.
// mode 0 (epipolar) needs the two-view geometry, which the reference's C++ never computes (its EpiSGM::compute ignores K as
// well): -g names a text file with 9 numbers F (row-major), 9 numbers H, 2 numbers epipole (1-based), 1 number direction.
#include "../../include/fsgm_proj.hpp"
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <map>
#include <stdexcept>
#include <string>

using namespace fsgm_proj;

static const char* kUsage =
    "SGM OF v0.0.1\nCall SGMOF to do optical flow for two images, Usage:\n"
    " ./sgmof I1<first image> I2<second image> [-o]=<output flow file name> [-m]=0(epiSGM)/1(pydSGM)\n"
    "  -o, --outFile         output flow file (in KITTI format)                     [flow.png]\n"
    "  -m, --mode            epiSGM(0)/pydSGM mode(1)                               [0]\n"
    "  -c, --calibFile       calibration file, must have when mode = 0              [calib.txt]\n"
    "  -b, --benchmark       0/1 for Kitti2012/kitti2015 calibration file format    [0]\n"
    "  -p, --passNum         number of SGM passes                                   [2]\n"
    "  -d, --enableDiagonal  enable diagonal directions in SGM\n"
    "  -V, --vzIndex         enable vz-index in epipolar SGM\n"
    "  -N, --pydNum          number of pyramidal levels in PydSGM                   [5]\n"
    "  -g, --geometry        F, H, epipole, direction for mode 0 (21 numbers; see the header of sgmof_main.cpp)\n"
    "  -G, --groundTruth     KITTI flow PNG to score the result against\n";

int main(int argc, char** argv)
{
    std::map<std::string, std::string> opt = {{"o", "flow.png"}, {"m", "0"}, {"c", "calib.txt"}, {"b", "0"}, {"p", "2"}, {"N", "5"}};
    const std::map<std::string, std::string> longnames = {{"outFile", "o"}, {"mode", "m"}, {"calibFile", "c"}, {"benchmark", "b"},
        {"passNum", "p"}, {"enableDiagonal", "d"}, {"vzIndex", "V"}, {"pydNum", "N"}, {"geometry", "g"}, {"groundTruth", "G"}, {"help", "h"}};
    std::string pos[2];
    int npos = 0;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        if (a.size() > 1 && a[0] == '-') {
            a = a.substr(a[1] == '-' ? 2 : 1);
            std::string val;
            const size_t eq = a.find('=');
            bool has_val = eq != std::string::npos;
            if (has_val) { val = a.substr(eq + 1); a = a.substr(0, eq); }
            if (longnames.count(a)) a = longnames.at(a);
            if (a == "h" || a == "?" || a == "usage") { std::fputs(kUsage, stdout); return 0; }
            const bool flag = a == "d" || a == "V";
            if (!has_val && !flag) {
                if (i + 1 >= argc) { std::fprintf(stderr, "option -%s needs a value\n", a.c_str()); return 1; }
                val = argv[++i];
            }
            opt[a] = flag && !has_val ? "1" : val;
        } else if (npos < 2) pos[npos++] = a;
    }
    if (npos < 2) { std::fputs(kUsage, stderr); return 1; }
    std::string err;
    const Image I1 = imread(pos[0], &err), I2 = imread(pos[1], &err);
    if (I1.empty() || I2.empty()) { std::printf("Open image failed...\n%s\n", err.c_str()); return 1; }
    if (I1.rows != I2.rows || I1.cols != I2.cols) { std::printf("Size of image1/2 must match\n"); return 1; }
    FlowField flow;
    try {
        if (std::atoi(opt["m"].c_str()) == 0) {
            float P[12];
            if (!read_calib_file(opt["c"], std::atoi(opt["b"].c_str()) == 1, P, &err)) { std::printf("%s\n", err.c_str()); return 1; }
            std::printf("K = [%g 0 %g; 0 %g %g; 0 0 1]\n", P[0], P[2], P[5], P[6]);
            if (!opt.count("g")) { std::printf("mode 0 needs -g <geometry file>\n"); return 1; }
            std::ifstream g(opt["g"]);
            double v[21];
            for (int i = 0; i < 21; ++i)
                if (!(g >> v[i])) { std::printf("can't read 21 numbers from %s\n", opt["g"].c_str()); return 1; }
            EpiSGM epi;
            epi.enableDiagonal = opt.count("d") && opt["d"] != "0";
            epi.vzIndex = true;
            epi.setGeometry(v, v + 9, v + 18, v[20] != 0.0);
            flow = epi.compute(I1, I2);
        } else {
            PydSGM pyd;
            pyd.numPyd = std::atoi(opt["N"].c_str());
            pyd.passNum = std::atoi(opt["p"].c_str());
            pyd.enableDiagonal = !opt.count("d") || opt["d"] != "0";       // pyramidal_sgm.m:19 runs with the diagonals on
            flow = pyd.compute(I1, I2);
        }
    } catch (const std::exception& e) {
        std::printf("%s\n", e.what());
        return 1;
    }
    if (!flow_write_kitti(opt["o"], flow, &err)) { std::printf("%s\n", err.c_str()); return 1; }
    std::printf("wrote %s (%d x %d)\n", opt["o"].c_str(), flow.cols, flow.rows);
    if (opt.count("G")) {
        FlowField gt;
        if (!flow_read_kitti(opt["G"], &gt, &err)) { std::printf("%s\n", err.c_str()); return 1; }
        double epe = 0;
        const double out = flow_outlier_rate(flow, gt, &epe);
        std::printf("KITTI outliers (>3 px and >5 %%): %.2f %%, mean EPE %.3f px\n", 100.0 * out, epe);
    }
    return 0;
}
