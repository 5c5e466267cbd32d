/* ORACLE INFRASTRUCTURE: the one OMPL symbol the reference test binary imports (logging). */
namespace ompl { namespace msg { enum LogLevel { A }; void log(const char*, int, LogLevel, const char*, ...) {} } }
