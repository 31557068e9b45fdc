// TEST INFRASTRUCTURE ONLY (oracle build). boost::variate_generator<Engine, Distribution>
// with Engine held BY VALUE, as the reference instantiates it (Lsh.cpp:79).
#ifndef EM2_ORACLE_SHIM_VARIATE_HPP
#define EM2_ORACLE_SHIM_VARIATE_HPP
namespace boost {
template <class Engine, class Distribution> class variate_generator {
public:
    typedef typename Distribution::result_type result_type;
    variate_generator(Engine e, Distribution d) : eng_(e), dist_(d) {}
    result_type operator()() { return dist_(eng_); }
    Engine& engine() { return eng_; }
    Distribution& distribution() { return dist_; }
private:
    Engine eng_;
    Distribution dist_;
};
}
#endif
